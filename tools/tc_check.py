#!/usr/bin/env python
"""Diagnostic run of the tensor-core PHMLinear path against the fp64 oracle (prints, never asserts)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import phc_oracle as O
from phc_gnn_b200 import ops

DEV = "cuda:0"


def case(n, fin, fout, M, precision, bwd=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(n, n, n, generator=g) * 0.5
    W = torch.randn(n, fin // n, fout // n, generator=g) * 0.2
    b = torch.randn(fout, generator=g)
    x = torch.randn(M, fin, generator=g)
    res = torch.randn(M, fout, generator=g)
    gy = torch.randn(M, fout, generator=g)
    po = {"l.phm_rule": A.double().requires_grad_(True), "l.W": W.double().requires_grad_(True), "l.b": b.double().requires_grad_(True)}
    xo = x.double().requires_grad_(True)
    ref = O.phm_linear(xo, po, "l") + res.double()
    ref.backward(gy.double())
    t = [v.to(DEV).requires_grad_(True) for v in (x, A, W, b, res)]
    torch.cuda.synchronize()
    y = ops.phm_linear(t[0], t[1], t[2], t[3], t[4], precision=precision)
    torch.cuda.synchronize()
    def rel(a, r):
        return float((a.detach().cpu().double() - r.detach()).abs().max() / r.detach().abs().max())
    msg = f"n={n} in={fin} out={fout} M={M} prec={precision}: y {rel(y, ref):.2e}"
    if bwd:
        y.backward(gy.to(DEV))
        torch.cuda.synchronize()
        msg += f" dx {rel(t[0].grad, xo.grad):.2e} dA {rel(t[1].grad, po['l.phm_rule'].grad):.2e} dW {rel(t[2].grad, po['l.W'].grad):.2e} db {rel(t[3].grad, po['l.b'].grad):.2e}"
    print(msg, flush=True)


if __name__ == "__main__":
    bwd = "--fwd-only" not in sys.argv
    for prec in (1,):
        for (n, fin, fout, M) in ((4, 128, 128, 512), (4, 32, 32, 512), (4, 500, 500, 1000), (2, 180, 180, 3000), (4, 200, 200, 3333),
                                  (1, 224, 56, 9000), (5, 200, 200, 777), (8, 512, 512, 640), (4, 512, 768, 600), (3, 33, 300, 515),
                                  (16, 512, 64, 520)):
            case(n, fin, fout, M, prec, bwd)
    # timing
    for (n, F, M) in ((4, 500, 15600), (4, 512, 13300), (4, 200, 3300)):
        x = torch.randn(M, F, device=DEV, requires_grad=True); A = torch.randn(n, n, n, device=DEV, requires_grad=True)
        W = torch.randn(n, F // n, F // n, device=DEV, requires_grad=True); b = torch.randn(F, device=DEV, requires_grad=True)
        gy = torch.randn(M, F, device=DEV)
        for prec in (0, 1, 2):
            for _ in range(3):
                y = ops.phm_linear(x, A, W, b, precision=prec); y.backward(gy)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            for _ in range(10):
                y = ops.phm_linear(x, A, W, b, precision=prec)
            e[1].record()
            for _ in range(10):
                y.backward(gy, retain_graph=True)
            e[2].record()
            torch.cuda.synchronize()
            f, bw = e[0].elapsed_time(e[1]) / 10, e[1].elapsed_time(e[2]) / 10
            fl = 2.0 * M * F * F
            print(f"timing n={n} F={F} M={M} prec={prec}: fwd {f*1e3:.1f} us ({fl/f/1e9:.1f} TFLOP/s)  bwd {bw*1e3:.1f} us ({2*fl/bw/1e9:.1f} TFLOP/s)", flush=True)
