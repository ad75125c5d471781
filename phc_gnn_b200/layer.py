"""One autograd node per message-passing layer.

``PHMSkipConnectAdd`` spends one layer as  conv (aggregate + edge encoder) -> PHM transform (2-layer PHM MLP with
batch-norm, or a single PHMLinear with residual) -> batch-norm -> activation -> dropout -> skip add
(reference models.py:200-217, messagepassing.py:55-70,132-142, layers.py:349-355).  Issuing that as five
separate autograd Functions and ten nn.Module calls costs more host time than the kernels take on the GPU for
molecule-sized batches, so the model's fast path runs the whole layer inside ONE autograd Function: forward
launches the five kernels back to back, backward launches their gradients in reverse.  The arithmetic and the
kernels are exactly those of ops.py; only the host-side orchestration differs.

Two implementations of the node: ``_ConvLayer`` issues one C-ABI call per operator (used when the per-operator
CUDA-event instrumentation is on, and as the cross-check in the tests); ``_ConvLayerCall`` hands the whole layer to
the library in ONE call per direction (``phc_conv_layer_fwd`` / ``phc_conv_layer_bwd``, include/phc_b200_layer.h) —
at ppa shape the Python cost of thirteen foreign calls per layer had caught up with the GPU time of their kernels.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib
from .graph import EdgeStructure, _stream
from .parallel import grad_sink
from .ops import PROFILE, REDUCE_IDS, _ptr, _ptr_array, _ws, act_id, default_precision, next_dropout_seed, run

_WS_CACHE = {}


def _ws_bytes(name: str, *args) -> int:
    key = (name,) + args
    v = _WS_CACHE.get(key)
    if v is None:
        v = getattr(_lib.load(), name)(*args)
        _WS_CACHE[key] = v
    return v


def _lin_fwd(x, rule, W, b, residual, precision, stream):
    n, K, P = W.shape
    M = x.size(0)
    y = torch.empty((M, n * P), dtype=torch.float32, device=x.device)
    nb = _ws_bytes("phc_phm_linear_fwd_workspace_bytes", M, n * K, n * P, n, precision)
    ws = _ws(nb, x.device)
    run("phc_phm_linear_fwd", None, x.data_ptr(), rule.data_ptr(), W.data_ptr(), _ptr(b), _ptr(residual), y.data_ptr(), M, n * K, n * P,
        n, 0, precision, ws.data_ptr(), ws.numel(), stream, tag=":node" if M >= 1024 else ":head")
    return y, (ws if nb > 64 else None)


def _lin_bwd(gy, x, rule, W, has_bias, need_dx, precision, fwd_ws, stream):
    n, K, P = W.shape
    M = x.size(0)
    dx = torch.empty_like(x) if need_dx else None
    d_rule = torch.empty_like(rule) if rule.requires_grad else None
    dW = torch.empty_like(W)
    db = torch.empty(n * P, dtype=torch.float32, device=x.device) if has_bias else None
    nb = _ws_bytes("phc_phm_linear_bwd_workspace_bytes", M, n * K, n * P, n, precision)
    ws = _ws(nb, x.device)
    run("phc_phm_linear_bwd", None, gy.data_ptr(), x.data_ptr(), rule.data_ptr(), W.data_ptr(), _ptr(dx), _ptr(d_rule), dW.data_ptr(),
        _ptr(db), M, n * K, n * P, n, precision, ws.data_ptr(), ws.numel(), _ptr(fwd_ws), stream, tag=":node" if M >= 1024 else ":head")
    return dx, d_rule, dW, db


def _bn_fwd(h, flat, skip, n, use_bn, training, momentum, eps, act, p, same, seed, stream):
    M, F = h.shape
    gamma, beta, rmean, rvar, tracked = flat if flat is not None else (None,) * 5
    y = torch.empty_like(h)
    stats = torch.empty((2, F), dtype=torch.float32, device=h.device) if use_bn else None
    nb = _ws_bytes("phc_bn_workspace_bytes", M, F) if (use_bn and training) else 16
    ws = _ws(nb, h.device)
    upd = training and use_bn
    run("phc_bn_act_drop_skip_fwd", None, h.data_ptr(), _ptr(gamma), _ptr(beta), _ptr(rmean) if (upd or not training) else 0,
        _ptr(rvar) if (upd or not training) else 0, _ptr(tracked) if upd else 0, 0 if tracked is None else tracked.numel(), _ptr(skip),
        M, F, n, int(use_bn), int(training), momentum, eps, act, float(p), int(same), seed, y.data_ptr(),
        _ptr(stats[0]) if use_bn else 0, _ptr(stats[1]) if use_bn else 0, ws.data_ptr(), ws.numel(), stream)
    return y, stats


def _bn_bwd(gy, h, flat, stats, n, use_bn, training, act, p, same, seed, stream):
    M, F = h.shape
    gamma, beta = (flat[0], flat[1]) if flat is not None else (None, None)
    dh = torch.empty_like(h)
    dgb = torch.empty((2, F), dtype=torch.float32, device=h.device) if use_bn else None
    nb = _ws_bytes("phc_bn_workspace_bytes", M, F) if use_bn else 16
    ws = _ws(nb, h.device)
    run("phc_bn_act_drop_skip_bwd", None, gy.data_ptr(), h.data_ptr(), _ptr(gamma), _ptr(beta), _ptr(stats[0]) if use_bn else 0,
        _ptr(stats[1]) if use_bn else 0, M, F, n, int(use_bn), int(training), act, float(p), int(same), seed, dh.data_ptr(),
        _ptr(dgb[0]) if use_bn else 0, _ptr(dgb[1]) if use_bn else 0, ws.data_ptr(), ws.numel(), stream)
    return dh, dgb


def _split_gb(dgb, count, F):
    """[2,F] (dgamma; dbeta) -> count views of dgamma followed by count views of dbeta."""
    if dgb is None or count == 0:
        return []
    fc = F // count
    return [dgb[0, c * fc:(c + 1) * fc] for c in range(count)] + [dgb[1, c * fc:(c + 1) * fc] for c in range(count)]


class _ConvLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, struct: EdgeStructure, flats, x, skip, attr, *tensors):
        (n, linear, enc_dim, vocab, reduce, msg_act, self_loops, mlp, act1, act2, use_bn1, use_bn2, training, drop_p, same, seed,
         precision, n_enc, has_beta, nb1, nb2, mom1, eps1, mom2, eps2) = cfg
        flat1, flat2 = flats
        dev = x.device
        st = _stream(dev)
        i = 0
        beta = tensors[0] if has_beta else None
        i += 1 if has_beta else 0
        enc = tensors[i:i + n_enc]; i += n_enc
        r1, W1, b1 = tensors[i:i + 3]; i += 3
        i += nb1
        if mlp:
            r2, W2, b2 = tensors[i:i + 3]; i += 3
        N, F = x.shape
        # 1. aggregation with the edge encoder fused in
        agg = torch.empty_like(x)
        aux_f = torch.empty((2, N, F), dtype=torch.float32, device=dev) if reduce == 4 else None
        aux_i = torch.empty((N, F), dtype=torch.int32, device=dev) if reduce in (2, 3) else None
        vc = (ctypes.c_int * max(len(vocab), 1))(*vocab) if vocab else None
        rows = enc_dim + 1 if linear else int(sum(vocab))
        sums = _node_sums(struct, attr, linear, enc_dim, vocab, vc, reduce, msg_act, rows, N, dev, st) if NODE_SUM_FORWARD else None
        if sums is not None:
            tws = _ws(_ws_bytes("phc_conv_fused_fwd_sums_workspace_bytes", F, rows), dev)
            run("phc_conv_fused_fwd_sums", None, x.data_ptr(), sums.data_ptr(), 0 if linear else 1, enc_dim, vc, _ptr_array(enc),
                struct.rowptr.data_ptr(), struct.col.data_ptr(), N, F, n, reduce, int(self_loops and mlp), agg.data_ptr(),
                tws.data_ptr(), tws.numel(), st, launches=2)
        else:
            run("phc_conv_fused_fwd", None, x.data_ptr(), attr.data_ptr(), 0 if linear else 1, enc_dim, vc, _ptr_array(enc),
                struct.rowptr.data_ptr(), struct.col.data_ptr(), struct.perm.data_ptr(), N, F, n, reduce, msg_act, _ptr(beta),
                int(self_loops and mlp), agg.data_ptr(), _ptr(aux_f), _ptr(aux_i), st)
        # 2.-4. PHM transform
        if mlp:
            y1, ws1 = _lin_fwd(agg, r1, W1, b1, None, precision, st)
            a1, stats1 = _bn_fwd(y1, flat1, None, n, use_bn1, training, mom1, eps1, act1, 0.0, False, 0, st)
            z, ws2 = _lin_fwd(a1, r2, W2, b2, None, precision, st)
        else:
            y1 = a1 = stats1 = ws2 = None
            z, ws1 = _lin_fwd(agg, r1, W1, b1, x if self_loops else None, precision, st)
        # 5. norm -> act -> dropout -> + skip
        out, stats2 = _bn_fwd(z, flat2, skip, n, use_bn2, training, mom2, eps2, act2, drop_p, same, seed, st)
        ctx.save_for_backward(x, attr, agg, y1, a1, z, aux_f, aux_i, stats1, stats2, *tensors)
        ctx.misc = (cfg, struct, flats, ws1, ws2)
        ctx.has_skip = skip is not None
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        cfg, struct, flats, ws1, ws2 = ctx.misc
        (n, linear, enc_dim, vocab, reduce, msg_act, self_loops, mlp, act1, act2, use_bn1, use_bn2, training, drop_p, same, seed,
         precision, n_enc, has_beta, nb1, nb2, mom1, eps1, mom2, eps2) = cfg
        flat1, flat2 = flats
        x, attr, agg, y1, a1, z, aux_f, aux_i, stats1, stats2 = ctx.saved_tensors[:10]
        tensors = ctx.saved_tensors[10:]
        i = 0
        beta = tensors[0] if has_beta else None
        i += 1 if has_beta else 0
        enc = tensors[i:i + n_enc]; i += n_enc
        r1, W1, b1 = tensors[i:i + 3]; i += 3 + nb1
        if mlp:
            r2, W2, b2 = tensors[i:i + 3]
        dev = x.device
        st = _stream(dev)
        g = g.contiguous()
        N, F = x.shape
        # 5'
        dz, dgb2 = _bn_bwd(g, z, flat2, stats2, n, use_bn2, training, act2, drop_p, same, seed, st)
        # 4'-2'
        if mlp:
            da1, dr2, dW2, db2 = _lin_bwd(dz, a1, r2, W2, b2 is not None, True, precision, ws2, st)
            dy1, dgb1 = _bn_bwd(da1, y1, flat1, stats1, n, use_bn1, training, act1, 0.0, False, 0, st)
            dagg, dr1, dW1, db1 = _lin_bwd(dy1, agg, r1, W1, b1 is not None, True, precision, ws1, st)
        else:
            dgb1 = None
            dagg, dr1, dW1, db1 = _lin_bwd(dz, agg, r1, W1, b1 is not None, True, precision, ws1, st)
        # 1'
        dx = torch.empty_like(x)
        flatg = torch.empty(sum(p.numel() for p in enc), dtype=torch.float32, device=dev)
        genc, o = [], 0
        for p in enc:
            genc.append(flatg[o:o + p.numel()].view(p.shape))
            o += p.numel()
        dbeta = torch.zeros((), dtype=torch.float32, device=dev) if reduce == 4 else None
        rows = enc_dim + 1 if linear else int(sum(vocab))
        nb = _ws_bytes("phc_conv_fused_bwd_workspace_bytes", N, F, rows)
        ws = _ws(nb, dev)
        vc = (ctypes.c_int * max(len(vocab), 1))(*vocab) if vocab else None
        sums = None
        if reduce in (0, 1) and msg_act == 0 and rows <= 16:
            key = ("sums", attr.data_ptr(), tuple(attr.shape), attr._version, linear, enc_dim, vocab, reduce == 1)
            sums = struct.extras.get(key)
            if sums is None:
                sums = torch.empty((N, rows), dtype=torch.float32, device=dev)
                run("phc_edge_feature_sums", None, attr.data_ptr(), 0 if linear else 1, enc_dim, vc, struct.rowptr.data_ptr(),
                    struct.perm.data_ptr(), N, int(reduce == 1), sums.data_ptr(), st)
                struct.extras[key] = sums
                struct.extras[("keepalive", attr.data_ptr())] = attr
        run("phc_conv_fused_bwd", None, dagg.data_ptr(), x.data_ptr(), attr.data_ptr(), 0 if linear else 1, enc_dim, vc, _ptr_array(enc),
            _ptr_array(genc), _ptr(aux_f), _ptr(aux_i), struct.rowptr.data_ptr(), struct.col.data_ptr(), struct.perm.data_ptr(),
            struct.rowptr_t.data_ptr(), struct.col_t.data_ptr(), struct.perm_t.data_ptr(), N, F, n, reduce, msg_act, _ptr(beta),
            int(self_loops and mlp), _ptr(sums), dx.data_ptr(), _ptr(dbeta), ws.data_ptr(), ws.numel(), st)
        if not mlp and self_loops:
            dx.add_(dz)                         # residual branch of PHMLinear(agg) + x
        grads = []
        if has_beta:
            grads.append(dbeta)
        grads += genc
        grads += [dr1, dW1, db1]
        grads += _split_gb(dgb1, nb1 // 2, F) if nb1 else []
        if mlp:
            grads += [dr2, dW2, db2]
        grads += _split_gb(dgb2, nb2 // 2, F) if nb2 else []
        return (None, None, None, dx, g if ctx.has_skip else None, None) + tuple(grads)


SINGLE_CALL = True      # False: one C-ABI call per operator (the cross-check path of the tests)
DIRECT_PARAM_GRADS = True   # parameter gradients written in place by the layer's backward instead of flowing through autograd
NODE_SUM_FORWARD = True     # sum / mean + identity message: encoder term from the per-node feature sums (pure row gather)
_SCRATCH = {}


def _scratch(nbytes: int, device, stream: int) -> torch.Tensor:
    """Grow-only scratch per (device, stream): every user is ordered on that stream, so one buffer serves all layers."""
    key = (device.index, stream)
    t = _SCRATCH.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _SCRATCH[key] = t
    return t


def _node_sums(struct, attr, linear, enc_dim, vocab, vc, reduce, msg_act, rows, N, dev, st):
    """Per-node sums of the raw edge features (depends only on the batch: shared by all layers through struct.extras)."""
    if not (reduce in (0, 1) and msg_act == 0 and rows <= 16):
        return None
    key = ("sums", attr.data_ptr(), tuple(attr.shape), attr._version, linear, enc_dim, vocab, reduce == 1)
    sums = struct.extras.get(key)
    if sums is None:
        sums = torch.empty((N, rows), dtype=torch.float32, device=dev)
        run("phc_edge_feature_sums", None, attr.data_ptr(), 0 if linear else 1, enc_dim, vc, struct.rowptr.data_ptr(),
            struct.perm.data_ptr(), N, int(reduce == 1), sums.data_ptr(), st)
        struct.extras[key] = sums
        struct.extras[("keepalive", attr.data_ptr())] = attr
    return sums


def _unpack(cfg, tensors):
    """tensors = [beta?] + encoder params + [rule1, W1, b1] + bn1 params + [rule2, W2, b2]? + bn2 params (see conv_layer)."""
    n_enc, has_beta, nb1, mlp = cfg[17], cfg[18], cfg[19], cfg[7]
    i = 1 if has_beta else 0
    beta = tensors[0] if has_beta else None
    enc = tensors[i:i + n_enc]; i += n_enc
    lin1 = tensors[i:i + 3]; i += 3
    bn1 = tensors[i:i + nb1]; i += nb1
    lin2 = (None, None, None)
    if mlp:
        lin2 = tensors[i:i + 3]; i += 3
    bn2 = tensors[i:]
    return beta, enc, lin1, bn1, lin2, bn2


def _call_fwd(cfg, struct: EdgeStructure, flats, x, skip, attr, tensors):
    """Whole layer forward through phc_conv_layer_fwd.  -> (out, saved tensors, host-side state for backward)."""
    (n, linear, enc_dim, vocab, reduce, msg_act, self_loops, mlp, act1, act2, use_bn1, use_bn2, training, drop_p, same, seed,
     precision, n_enc, has_beta, nb1, nb2, mom1, eps1, mom2, eps2) = cfg
    flat1, flat2 = flats
    dev = x.device
    st = _stream(dev)
    beta, enc, (r1, W1, b1), _, (r2, W2, b2), _ = _unpack(cfg, tensors)
    N, F = x.shape
    f32 = dict(dtype=torch.float32, device=dev)
    acts = torch.empty((4 if mlp else 2, N, F), **f32)          # agg, [y1, a1,] z: saved for backward
    out = torch.empty((N, F), **f32)                            # own storage: callers may modify it in place
    stats = torch.empty((4, F), **f32)
    aux_f = torch.empty((2, N, F), **f32) if reduce == 4 else None
    aux_i = torch.empty((N, F), dtype=torch.int32, device=dev) if reduce in (2, 3) else None
    nb_lin = _ws_bytes("phc_phm_linear_fwd_workspace_bytes", N, F, F, n, precision)
    nb_lin = (nb_lin + 1023) & ~1023
    ws_lin = _ws(nb_lin * (2 if mlp else 1), dev)               # PHMLinear operand packs, kept for backward
    rows = enc_dim + 1 if linear else int(sum(vocab))
    ws = _scratch(_ws_bytes("phc_conv_layer_workspace_bytes", N, F, n, rows, precision), dev, st)
    vc = (ctypes.c_int * max(len(vocab), 1))(*vocab) if vocab else None
    encp = _ptr_array(enc)
    D = _lib.conv_layer_struct()()
    D.num_nodes, D.width, D.phm_dim = N, F, n
    D.enc_kind, D.enc_dim, D.table_rows = 0 if linear else 1, enc_dim, rows
    D.reduce, D.msg_act, D.self_loops, D.mlp = reduce, msg_act, int(self_loops), int(mlp)
    D.act1, D.act2, D.use_bn1, D.use_bn2 = act1, act2, int(use_bn1 and mlp), int(use_bn2)
    D.training, D.drop_same, D.precision = int(training), int(same), precision
    D.drop_p, D.momentum1, D.eps1, D.momentum2, D.eps2, D.seed = float(drop_p), mom1, eps1, mom2, eps2, seed
    D.rowptr, D.col, D.perm = struct.rowptr.data_ptr(), struct.col.data_ptr(), struct.perm.data_ptr()
    D.rowptr_t, D.col_t, D.perm_t = struct.rowptr_t.data_ptr(), struct.col_t.data_ptr(), struct.perm_t.data_ptr()
    D.x, D.skip, D.edge_attr = x.data_ptr(), _ptr(skip), attr.data_ptr()
    D.vocab = ctypes.cast(vc, ctypes.c_void_p) if vc is not None else None
    D.enc_params = ctypes.cast(encp, ctypes.c_void_p)
    D.softmax_beta = _ptr(beta)
    D.rule1, D.W1, D.b1 = r1.data_ptr(), W1.data_ptr(), _ptr(b1)
    if mlp:
        D.rule2, D.W2, D.b2 = r2.data_ptr(), W2.data_ptr(), _ptr(b2)
    if flat1 is not None and mlp:
        gamma, bt, rmean, rvar, tracked = flat1
        D.gamma1, D.beta1, D.running_mean1, D.running_var1 = _ptr(gamma), _ptr(bt), _ptr(rmean), _ptr(rvar)
        D.tracked1, D.n_tracked1 = _ptr(tracked), 0 if tracked is None else tracked.numel()
    if flat2 is not None:
        gamma, bt, rmean, rvar, tracked = flat2
        D.gamma2, D.beta2, D.running_mean2, D.running_var2 = _ptr(gamma), _ptr(bt), _ptr(rmean), _ptr(rvar)
        D.tracked2, D.n_tracked2 = _ptr(tracked), 0 if tracked is None else tracked.numel()
    ap = acts.data_ptr()
    sz = N * F * 4
    if mlp:
        D.agg, D.y1, D.a1, D.z = ap, ap + sz, ap + 2 * sz, ap + 3 * sz
    else:
        D.agg, D.z = ap, ap + sz
    D.out = out.data_ptr()
    D.stats1, D.stats2 = stats.data_ptr(), stats.data_ptr() + 2 * F * 4
    D.aux_f, D.aux_i = _ptr(aux_f), _ptr(aux_i)
    D.ws_lin1, D.ws_lin1_bytes = ws_lin.data_ptr(), nb_lin
    if mlp:
        D.ws_lin2, D.ws_lin2_bytes = ws_lin.data_ptr() + nb_lin, nb_lin
    D.ws, D.ws_bytes = ws.data_ptr(), ws.numel()
    sums = _node_sums(struct, attr, linear, enc_dim, vocab, vc, reduce, msg_act, rows, N, dev, st) if NODE_SUM_FORWARD else None
    D.node_sums = _ptr(sums)
    fused_stats = precision != 0 and n == 4 and training        # batch-norm statistics come out of the PHMLinear epilogue
    run("phc_conv_layer_fwd", None, ctypes.byref(D), st,
        launches=(11 if mlp else 6) - ((int(use_bn1 and mlp) + int(use_bn2)) if fused_stats else 0))
    return out, (acts, stats, aux_f, aux_i), (D, vc, encp, ws_lin)


def _call_bwd(cfg, struct: EdgeStructure, host, x, attr, g, tensors, sinks=None):
    """Whole layer backward through phc_conv_layer_bwd.  ``sinks`` (optional, aligned with ``tensors``): tensors to write
    the parameter gradients INTO (None entries get fresh storage).  -> (dx, grads aligned with ``tensors``)."""
    (n, linear, enc_dim, vocab, reduce, msg_act, self_loops, mlp, act1, act2, use_bn1, use_bn2, training, drop_p, same, seed,
     precision, n_enc, has_beta, nb1, nb2, mom1, eps1, mom2, eps2) = cfg
    D, vc = host[0], host[1]
    dev = x.device
    st = _stream(dev)
    g = g.contiguous()
    N, F = x.shape
    f32 = dict(dtype=torch.float32, device=dev)
    tmp = torch.empty((2, N, F), **f32)
    dx = torch.empty_like(x)
    nt = len(tensors)
    grads = [None] * nt
    if sinks is None:
        sinks = [None] * nt
    # which gradients are wanted, and how large is the fresh storage for those without a sink
    i0 = 1 if has_beta else 0
    i_lin1 = i0 + n_enc
    i_bn1 = i_lin1 + 3
    i_lin2 = i_bn1 + nb1
    i_bn2 = i_lin2 + (3 if mlp else 0)
    single = list(range(i0, i_bn1)) + (list(range(i_lin2, i_bn2)) if mlp else [])      # encoder tables, rule / W / b

    def wanted(k):
        t = tensors[k]
        if t is None:
            return False
        if k < i_lin1:
            return True                                 # encoder tables: the kernel writes all of them
        j = (k - i_lin1) if k < i_bn1 else (k - i_lin2)
        return True if j == 1 else (t.requires_grad if j == 0 else True)     # rule: only when learned; W, b: always

    need = 0
    for k in single:
        if wanted(k) and sinks[k] is None:
            need += (tensors[k].numel() + 3) & ~3
    groups = []                                         # batch-norm affine groups: one [2, F] block = n gammas then n betas
    for start, cnt in ((i_bn1, nb1), (i_bn2, nb2)):
        if cnt == 0:
            groups.append(None)
            continue
        sk = sinks[start:start + cnt]
        base = sk[0].data_ptr() if sk[0] is not None else 0
        ok, off = base != 0, 0
        if ok:
            for q in sk:
                if q is None or q.data_ptr() != base + off * 4:
                    ok = False
                    break
                off += q.numel()
        groups.append((start, cnt, ok, base))
        if not ok:
            need += 2 * F
    fresh = torch.empty(need + 4, **f32) if need else None
    o = 0
    ptrs = [0] * nt
    for k in single:
        t = tensors[k]
        if not wanted(k):
            continue
        if sinks[k] is not None:
            grads[k] = sinks[k]
        else:
            grads[k] = fresh[o:o + t.numel()].view(t.shape)
            o += (t.numel() + 3) & ~3
        ptrs[k] = grads[k].data_ptr()
    gb_ptr = [0, 0]
    for gi, grp in enumerate(groups):
        if grp is None:
            continue
        start, cnt, ok, base = grp
        if ok:
            gb_ptr[gi] = base
            for q in range(cnt):
                grads[start + q] = sinks[start + q]
        else:
            blk = fresh[o:o + 2 * F].view(2, F)
            o += 2 * F
            gb_ptr[gi] = blk.data_ptr()
            for q, v in enumerate(_split_gb(blk, cnt // 2, F)):
                grads[start + q] = v
    dbeta = None
    if has_beta:
        dbeta = sinks[0] if sinks[0] is not None else torch.empty((), **f32)
        dbeta.zero_()
        grads[0] = dbeta
    rows = enc_dim + 1 if linear else int(sum(vocab))
    ws = _scratch(_ws_bytes("phc_conv_layer_workspace_bytes", N, F, n, rows, precision), dev, st)
    sums = _node_sums(struct, attr, linear, enc_dim, vocab, vc, reduce, msg_act, rows, N, dev, st)
    gencp = (ctypes.c_void_p * max(n_enc, 1))(*[ptrs[i0 + q] for q in range(n_enc)])
    D.ws, D.ws_bytes = ws.data_ptr(), ws.numel()
    D.gout, D.node_sums = g.data_ptr(), _ptr(sums)
    D.tmp_a, D.tmp_b, D.dx = tmp.data_ptr(), tmp.data_ptr() + N * F * 4, dx.data_ptr()
    D.d_softmax_beta = _ptr(dbeta)
    D.d_enc_params = ctypes.cast(gencp, ctypes.c_void_p)
    D.d_rule1, D.d_W1, D.d_b1 = ptrs[i_lin1], ptrs[i_lin1 + 1], ptrs[i_lin1 + 2]
    if mlp:
        D.d_rule2, D.d_W2, D.d_b2 = ptrs[i_lin2], ptrs[i_lin2 + 1], ptrs[i_lin2 + 2]
    D.d_gb1, D.d_gb2 = gb_ptr[0], gb_ptr[1]
    run("phc_conv_layer_bwd", None, ctypes.byref(D), st, launches=17 if mlp else 10)
    if not mlp and self_loops:
        dx.add_(tmp[0])                     # residual branch of PHMLinear(agg) + x
    return dx, g, grads


class _ConvLayerCall(torch.autograd.Function):
    """Same node as _ConvLayer, one library call per direction; parameter gradients returned through autograd."""

    @staticmethod
    def forward(ctx, cfg, struct: EdgeStructure, flats, x, skip, attr, *tensors):
        out, saved, host = _call_fwd(cfg, struct, flats, x, skip, attr, tensors)
        ctx.save_for_backward(x, attr, *saved)
        ctx.misc = (cfg, struct, host, tensors, skip is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        cfg, struct, host, tensors, has_skip = ctx.misc
        x, attr = ctx.saved_tensors[:2]
        dx, g, grads = _call_bwd(cfg, struct, host, x, attr, g, tensors)
        return (None, None, None, dx, g if has_skip else None, None) + tuple(grads)


class _ConvLayerDirect(torch.autograd.Function):
    """The node with the parameters OUTSIDE the autograd graph: only x and skip are graph inputs; backward writes every
    parameter gradient straight into its final place — the parameter's slice of the flat gradient buffer when a
    GradientBucket registered one (``parallel.register_grad_sink``), fresh storage otherwise — and assigns / accumulates ``param.grad``
    itself, as AccumulateGrad would.  At ppa shape the 24 parameter inputs per layer cost more host time in
    Function.apply, the engine and AccumulateGrad than the layer's kernels take on the GPU.  Not visible to
    torch.autograd.grad(), parameter hooks or double backward: set ``layer.DIRECT_PARAM_GRADS = False`` for those."""

    @staticmethod
    def forward(ctx, plan, x, skip):
        cfg, struct, flats, attr, tensors = plan
        out, saved, host = _call_fwd(cfg, struct, flats, x, skip, attr, tensors)
        ctx.save_for_backward(x, attr, *saved)
        # the parameters are not saved tensors here (they are outside the graph): do autograd's version check by hand
        ctx.misc = (cfg, struct, host, tensors, skip is not None, tuple(-1 if t is None else t._version for t in tensors))
        return out

    @staticmethod
    def backward(ctx, g):
        cfg, struct, host, tensors, has_skip, versions = ctx.misc
        for t, v in zip(tensors, versions):
            if t is not None and t._version != v:
                raise RuntimeError("a parameter of a fused message-passing layer was modified in place between its forward and "
                                   "backward (the backward would read the new value); clone it or set layer.DIRECT_PARAM_GRADS = False")
        x, attr = ctx.saved_tensors[:2]
        sinks = [None if (t is None or t.grad is not None) else grad_sink(t) for t in tensors]
        dx, g, grads = _call_bwd(cfg, struct, host, x, attr, g, tensors, sinks)
        for t, gr in zip(tensors, grads):
            if gr is None or t is None or not t.requires_grad:
                continue
            if t.grad is None:
                t.grad = gr
            else:
                t.grad.add_(gr)
        return None, dx, (g if has_skip else None)


def _direct_eligible(tensors) -> bool:
    """The in-place gradient path bypasses autograd for the layer's parameters, so it is taken only when nothing can observe the
    difference: every trainable parameter has a slice of a flat gradient buffer registered (a GradientBucket / FlatClipAdam owns
    the step) and none carries a tensor hook or a post-accumulate-grad hook (DistributedDataParallel, gradient clipping hooks,
    user hooks).  Anything else takes _ConvLayerCall, whose parameter gradients flow through autograd."""
    for t in tensors:
        if t is None or not t.requires_grad:
            continue
        if grad_sink(t) is None or t._backward_hooks or getattr(t, "_post_accumulate_grad_hooks", None):
            return False
    return True


def conv_layer(x, skip, edge_attr, struct: EdgeStructure, *, phm_dim: int, enc_linear: bool, enc_params: Sequence[torch.Tensor],
               enc_vocab: Sequence[int], reduce: str, msg_act: str, beta: Optional[torch.Tensor], add_self_loops: bool, mlp: bool,
               lin1, lin2, norm1, norm2, act1: str, act2: str, training: bool, drop_p: float, drop_same: bool) -> torch.Tensor:
    """Whole message-passing layer as one autograd node.  lin1/lin2: PHMLinear modules (lin2 None unless mlp);
    norm1/norm2: NaivePHMNorm modules or None."""
    r = REDUCE_IDS[reduce]
    if edge_attr.dim() == 1:
        edge_attr = edge_attr.unsqueeze(1)
    edge_attr = edge_attr.contiguous()
    edge_attr = edge_attr.to(torch.float32) if enc_linear else (edge_attr if edge_attr.dtype == torch.int64 else edge_attr.to(torch.int64))
    x = x.contiguous()
    if skip is not None:
        skip = skip.contiguous()
    active_drop = training and drop_p > 0.0
    seed = next_dropout_seed(x.device) if active_drop else 0
    has_beta = r == 4
    tensors = []
    if has_beta:
        tensors.append(beta)
    tensors += list(enc_params)
    tensors += [lin1.phm_rule, lin1.W, lin1.b]
    p1 = norm1.autograd_params() if norm1 is not None else []
    tensors += p1
    if mlp:
        tensors += [lin2.phm_rule, lin2.W, lin2.b]
    p2 = norm2.autograd_params() if norm2 is not None else []
    tensors += p2
    flat1 = norm1.flat_views() if norm1 is not None else None
    flat2 = norm2.flat_views() if norm2 is not None else None
    tr1 = (norm1.training or not norm1.track_running_stats) if norm1 is not None else training
    tr2 = (norm2.training or not norm2.track_running_stats) if norm2 is not None else training
    assert tr1 == training and tr2 == training, "fused layer: batch-norm modules and the layer disagree on train/eval mode"
    cfg = (phm_dim, bool(enc_linear), int(edge_attr.size(1)), tuple(int(v) for v in enc_vocab), r, act_id(msg_act), bool(add_self_loops),
           bool(mlp), act_id(act1), act_id(act2), norm1 is not None, norm2 is not None, bool(training),
           float(drop_p) if active_drop else 0.0, bool(drop_same), seed, default_precision(), len(enc_params), has_beta, len(p1), len(p2),
           float(norm1.momentum) if norm1 is not None else 0.1, float(norm1.eps) if norm1 is not None else 1e-5,
           float(norm2.momentum) if norm2 is not None else 0.1, float(norm2.eps) if norm2 is not None else 1e-5)
    if SINGLE_CALL and not PROFILE.timing:
        if DIRECT_PARAM_GRADS and torch.is_grad_enabled() and x.requires_grad and _direct_eligible(tensors):
            return _ConvLayerDirect.apply((cfg, struct, (flat1, flat2), edge_attr, tensors), x, skip)
        return _ConvLayerCall.apply(cfg, struct, (flat1, flat2), x, skip, edge_attr, *tensors)
    return _ConvLayer.apply(cfg, struct, (flat1, flat2), x, skip, edge_attr, *tensors)       # per-operator calls (each one timed)
