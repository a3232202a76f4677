"""torch.autograd bindings of the C-ABI kernels (thin: allocate, call, save).

PyTorch is plumbing here (device memory, streams, autograd graph); every number is
produced by the kernels in csrc/.  CUDA tensors only -- CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .layout import GraphLayout, _p, _require_cuda, _stream

EMBED_NAMES = ("theta1a", "theta1b", "theta2", "theta3", "theta4")
HEAD_NAMES = ("theta5", "theta6", "theta7")


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    _require_cuda(t, name)
    if t.dtype != torch.float32:
        raise TypeError("simmodel_b200: %s must be float32 (got %s)" % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def _embed_struct(T, w):
    """w: 10 tensors theta1a.w,b 1b.w,b 2.w,b 3.w,b 4.w,b"""
    prm = _lib.EmbedParams()
    prm.node_dim, prm.emb_dim, prm.rounds = int(w[0].shape[1]), int(w[0].shape[0]), int(T)
    prm.flags = _lib.path_flag()
    (prm.theta1a_w, prm.theta1a_b, prm.theta1b_w, prm.theta1b_b, prm.theta2_w, prm.theta2_b,
     prm.theta3_w, prm.theta3_b, prm.theta4_w, prm.theta4_b) = [_p(t) for t in w]
    return prm


def _head_struct(norm_agg, w):
    """w: 6 tensors theta5.w,b 6.w,b 7.w,b"""
    prm = _lib.HeadParams()
    prm.emb_dim, prm.norm_agg = int(w[2].shape[0]), int(bool(norm_agg))
    prm.flags = _lib.path_flag()
    prm.theta5_w, prm.theta5_b, prm.theta6_w, prm.theta6_b, prm.theta7_w, prm.theta7_b = [_p(t) for t in w]
    return prm


def _split_embed_grad(flat, F, P):
    sizes = [P * F, P, P * P, P, P * P, P, P * P, P, P, P]
    shapes = [(P, F), (P,), (P, P), (P,), (P, P), (P,), (P, P), (P,), (P, 1), (P,)]
    return [c.view(s) for c, s in zip(flat.split(sizes), shapes)]


def _split_head_grad(flat, P):
    sizes = [2 * P, 1, P * P, P, P * P, P]
    shapes = [(1, 2 * P), (1,), (P, P), (P,), (P, P), (P,)]
    return [c.view(s) for c, s in zip(flat.split(sizes), shapes)]


def _embed_forward(layout, T, x, w):
    L = _lib.lib()
    n, P = layout.num_nodes, int(w[0].shape[0])
    if x.shape[0] != n or x.shape[1] != w[0].shape[1]:
        raise ValueError("x has shape %s, expected (%d, %d)" % (tuple(x.shape), n, w[0].shape[1]))
    base = torch.empty(n, P, dtype=torch.float32, device=x.device)
    mus = torch.empty(T, n, P, dtype=torch.float32, device=x.device)
    prm = _embed_struct(T, w)
    with torch.cuda.device(x.device):
        _lib.check(L.s2v_embed_forward(C.byref(layout.c_struct()), C.byref(prm), _p(x), _p(base), _p(mus), _stream()),
                   "s2v_embed_forward")
    _lib.launch_count += T
    return base, mus


def qnet_small_eligible(layout, w) -> bool:
    """True when the single-launch small-graph kernel (csrc/s2v_small.cu) applies: the problem runs the CUDA-core path
    and is small enough that launches, not bytes, are the cost (acting, Manhattan5 / MemTask batches, PPO rollouts)."""
    return bool(_lib.lib().s2v_qnet_small_eligible(int(w[0].shape[0]), _lib.path_flag(), layout.num_nodes, layout.num_graphs))


def _qnet_small_forward(layout, T, x, ew, hw, norm_agg, mode, mask_u8=None, reach_ptr=None, reach_idx=None):
    """Whole forward pass in ONE cooperative launch.  mode None: embedding only.  Returns
    (base, mus, pool, gstate, q, prob, arg, node, val) -- entries a mode does not produce are None."""
    L = _lib.lib()
    n, B, P = layout.num_nodes, layout.num_graphs, int(ew[0].shape[0])
    dev = x.device
    if x.shape[0] != n or x.shape[1] != ew[0].shape[1]:
        raise ValueError("x has shape %s, expected (%d, %d)" % (tuple(x.shape), n, ew[0].shape[1]))
    base = torch.empty(n, P, dtype=torch.float32, device=dev)
    mus = torch.empty(T, n, P, dtype=torch.float32, device=dev)
    head = mode is not None
    pool = torch.empty(B, P, dtype=torch.float32, device=dev) if head else None
    gstate = torch.empty(B, P, dtype=torch.float32, device=dev) if head else None
    q = torch.empty(n, dtype=torch.float32, device=dev) if head else None
    prob = torch.empty(n, dtype=torch.float32, device=dev) if mode == "softmax" else None
    arg = torch.empty(B, dtype=torch.int64, device=dev) if mode in ("max", "list") else None
    node = torch.empty(B, dtype=torch.int64, device=dev) if mode == "list" else None
    val = torch.empty(B, dtype=torch.float32, device=dev) if mode in ("max", "list") else None
    eprm = _embed_struct(T, ew)
    hprm = _head_struct(norm_agg, hw) if head else None
    ws_floats = int(L.s2v_qnet_small_workspace_floats(P, n, B)) if head else 0       # many-CTA pool of large graphs
    ws = torch.empty(ws_floats, dtype=torch.float32, device=dev) if ws_floats else None
    with torch.cuda.device(dev):
        _lib.check(L.s2v_qnet_small_forward_ws(C.byref(layout.c_struct()), C.byref(eprm), C.byref(hprm) if head else None, _p(x),
                                               _p(base), _p(mus), _FUSED_MODES[mode] if head else -1, _p(mask_u8),
                                               _p(reach_ptr), _p(reach_idx), _p(pool), _p(gstate), _p(q), _p(prob), _p(arg),
                                               _p(node), _p(val), _p(ws), ws_floats, _stream()), "s2v_qnet_small_forward_ws")
    _lib.launch_count += 1
    return base, mus, pool, gstate, q, prob, arg, node, val


def _embed_backward(layout, T, x, w, base, mus, dmu, dmu_masked, dus=None, grad_out=None, want_dx=False):
    L = _lib.lib()
    n, P, F = layout.num_nodes, int(w[0].shape[0]), int(w[0].shape[1])
    if dus is None:
        dus = torch.empty(T, n, P, dtype=torch.float32, device=x.device)
    ws = torch.empty(L.s2v_embed_backward_workspace_bytes(F, P), dtype=torch.uint8, device=x.device)
    grad = grad_out if grad_out is not None else torch.empty(L.s2v_embed_grad_floats(F, P), dtype=torch.float32, device=x.device)
    prm = _embed_struct(T, w)
    dx = torch.empty(n, F, dtype=torch.float32, device=x.device) if want_dx else None
    with torch.cuda.device(x.device):
        _lib.check(L.s2v_embed_backward_dx(C.byref(layout.c_struct()), C.byref(prm), _p(x), _p(base), _p(mus), _p(dmu),
                                           int(dmu_masked), _p(dus), _p(ws), ws.numel(), _p(grad), _p(dx), _stream()),
                   "s2v_embed_backward")
    _lib.launch_count += T + 3 + (0 if dmu_masked else 1)
    if want_dx:
        return _split_embed_grad(grad, F, P), grad, dx
    return _split_embed_grad(grad, F, P), grad


def _head_forward(layout, norm_agg, mu, w):
    L = _lib.lib()
    n, B, P = layout.num_nodes, layout.num_graphs, int(w[2].shape[0])
    pool = torch.empty(B, P, dtype=torch.float32, device=mu.device)
    gstate = torch.empty(B, P, dtype=torch.float32, device=mu.device)
    q = torch.empty(n, dtype=torch.float32, device=mu.device)
    prm = _head_struct(norm_agg, w)
    with torch.cuda.device(mu.device):
        _lib.check(L.s2v_head_forward(C.byref(layout.c_struct()), C.byref(prm), _p(mu), _p(pool), _p(gstate), _p(q),
                                      _stream()), "s2v_head_forward")
    _lib.launch_count += 2
    return pool, gstate, q


def head_fused_eligible(layout, w) -> bool:
    """True when the fused head launch (pool + head + mask + reduction, csrc/s2v_head.cu:k_head_fused) computes the
    same bits as the three-launch sequence: the problem runs the CUDA-core head (small batches, acting)."""
    return bool(_lib.lib().s2v_head_fused_eligible(int(w[2].shape[0]), _lib.path_flag(), layout.num_nodes, layout.num_graphs))


_FUSED_MODES = {"q": 0, "softmax": 1, "max": 2, "list": 3}


def _head_fused_forward(layout, norm_agg, mu, w, mode, mask_u8=None, reach_ptr=None, reach_idx=None):
    """ONE launch: pool, gstate, q and -- per mode -- prob | (arg, val) | (pos, node, val)."""
    L = _lib.lib()
    n, B, P = layout.num_nodes, layout.num_graphs, int(w[2].shape[0])
    dev = mu.device
    pool = torch.empty(B, P, dtype=torch.float32, device=dev)
    gstate = torch.empty(B, P, dtype=torch.float32, device=dev)
    q = torch.empty(n, dtype=torch.float32, device=dev)
    prob = torch.empty(n, dtype=torch.float32, device=dev) if mode == "softmax" else None
    arg = torch.empty(B, dtype=torch.int64, device=dev) if mode in ("max", "list") else None
    node = torch.empty(B, dtype=torch.int64, device=dev) if mode == "list" else None
    val = torch.empty(B, dtype=torch.float32, device=dev) if mode in ("max", "list") else None
    prm = _head_struct(norm_agg, w)
    with torch.cuda.device(dev):
        _lib.check(L.s2v_head_fused_forward(C.byref(layout.c_struct()), C.byref(prm), _p(mu), _FUSED_MODES[mode], _p(mask_u8),
                                            _p(reach_ptr), _p(reach_idx), _p(pool), _p(gstate), _p(q), _p(prob), _p(arg),
                                            _p(node), _p(val), _stream()), "s2v_head_fused_forward")
    _lib.launch_count += 1
    return pool, gstate, q, prob, arg, node, val


def _head_backward(layout, norm_agg, mu, w, pool, gstate, dq, mask_relu, dmu_out=None, grad_out=None):
    L = _lib.lib()
    n, B, P = layout.num_nodes, layout.num_graphs, int(w[2].shape[0])
    if dmu_out is None:
        dmu_out = torch.empty(n, P, dtype=torch.float32, device=mu.device)
    ws = torch.empty(L.s2v_head_backward_workspace_bytes(P, B), dtype=torch.uint8, device=mu.device)
    grad = grad_out if grad_out is not None else torch.empty(L.s2v_head_grad_floats(P), dtype=torch.float32, device=mu.device)
    prm = _head_struct(norm_agg, w)
    with torch.cuda.device(mu.device):
        _lib.check(L.s2v_head_backward(C.byref(layout.c_struct()), C.byref(prm), _p(mu), _p(pool), _p(gstate), _p(dq),
                                       int(mask_relu), _p(dmu_out), _p(ws), ws.numel(), _p(grad), _stream()),
                   "s2v_head_backward")
    _lib.launch_count += 8 if n >= 4096 else 6       # + the streaming pass and the sparse-rows pass of large batches
    return dmu_out, _split_head_grad(grad, P), grad


class _EmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layout, T, x, *w):
        w = [_f32c(t.detach(), "parameter") for t in w]
        x = _f32c(x, "x")
        if qnet_small_eligible(layout, w):
            base, mus = _qnet_small_forward(layout, T, x, w, None, True, None)[:2]
        else:
            base, mus = _embed_forward(layout, T, x, w)
        ctx.layout, ctx.T = layout, T
        ctx.save_for_backward(x, base, mus, *w)
        return mus[T - 1]

    @staticmethod
    def backward(ctx, dmu):
        x, base, mus, *w = ctx.saved_tensors
        if ctx.needs_input_grad[2]:            # node features produced by a trainable module (lstm_type 'FE')
            grads, _, dx = _embed_backward(ctx.layout, ctx.T, x, w, base, mus, _f32c(dmu, "dmu"), False, want_dx=True)
            return (None, None, dx, *grads)
        grads, _ = _embed_backward(ctx.layout, ctx.T, x, w, base, mus, _f32c(dmu, "dmu"), False)
        return (None, None, None, *grads)


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layout, norm_agg, mu, *w):
        w = [_f32c(t.detach(), "parameter") for t in w]
        mu = _f32c(mu, "mu")
        pool, gstate, q = _head_forward(layout, norm_agg, mu, w)
        ctx.layout, ctx.norm_agg = layout, norm_agg
        ctx.save_for_backward(mu, pool, gstate, *w)
        return q

    @staticmethod
    def backward(ctx, dq):
        mu, pool, gstate, *w = ctx.saved_tensors
        dmu, grads, _ = _head_backward(ctx.layout, ctx.norm_agg, mu, w, pool, gstate, _f32c(dq, "dq"), False)
        return (None, None, dmu, *grads)


class _QNetFn(torch.autograd.Function):
    """Embedding + head in one node: the head backward hands du^T (already multiplied by
    [mu^T > 0]) straight to the embedding backward, written in place into the du workspace."""

    @staticmethod
    def forward(ctx, layout, T, norm_agg, x, *w):
        w = [_f32c(t.detach(), "parameter") for t in w]
        x = _f32c(x, "x")
        if qnet_small_eligible(layout, w[:10]):
            base, mus, pool, gstate, q = _qnet_small_forward(layout, T, x, w[:10], w[10:], norm_agg, "q")[:5]
        else:
            base, mus = _embed_forward(layout, T, x, w[:10])
            if head_fused_eligible(layout, w[10:]):
                pool, gstate, q = _head_fused_forward(layout, norm_agg, mus[T - 1], w[10:], "q")[:3]
            else:
                pool, gstate, q = _head_forward(layout, norm_agg, mus[T - 1], w[10:])
        ctx.layout, ctx.T, ctx.norm_agg = layout, T, norm_agg
        ctx.save_for_backward(x, base, mus, pool, gstate, *w)
        return q

    @staticmethod
    def backward(ctx, dq):
        x, base, mus, pool, gstate, *w = ctx.saved_tensors
        T, layout = ctx.T, ctx.layout
        n, P = layout.num_nodes, int(w[0].shape[0])
        dus = torch.empty(T, n, P, dtype=torch.float32, device=x.device)
        # ONE flat gradient buffer [embed | head] in QNet.parameters() order: the 16 .grad tensors autograd hands to the
        # parameters are views of it, so a data-parallel step all-reduces this buffer in place (dist.allreduce_gradients)
        L = _lib.lib()
        ne, nh = int(L.s2v_embed_grad_floats(int(w[0].shape[1]), P)), int(L.s2v_head_grad_floats(P))
        flat = torch.empty(ne + nh, dtype=torch.float32, device=x.device)
        _, hgrads, _ = _head_backward(layout, ctx.norm_agg, mus[T - 1], w[10:], pool, gstate, _f32c(dq, "dq"), True,
                                      dmu_out=dus[T - 1], grad_out=flat[ne:])
        if ctx.needs_input_grad[3]:
            egrads, _, dx = _embed_backward(layout, T, x, w[:10], base, mus, dus[T - 1], True, dus=dus, grad_out=flat[:ne],
                                            want_dx=True)
            return (None, None, None, dx, *egrads, *hgrads)
        egrads, _ = _embed_backward(layout, T, x, w[:10], base, mus, dus[T - 1], True, dus=dus, grad_out=flat[:ne])
        return (None, None, None, None, *egrads, *hgrads)


class _PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layout, norm_agg, mu):
        L = _lib.lib()
        mu = _f32c(mu, "mu")
        P = int(mu.shape[1])
        pool = torch.empty(layout.num_graphs, P, dtype=torch.float32, device=mu.device)
        with torch.cuda.device(mu.device):
            _lib.check(L.s2v_pool_forward(C.byref(layout.c_struct()), P, int(bool(norm_agg)), _p(mu), _p(pool), _stream()),
                       "s2v_pool_forward")
        _lib.launch_count += 1
        ctx.layout, ctx.norm_agg, ctx.P = layout, norm_agg, P
        return pool

    @staticmethod
    def backward(ctx, dpool):
        L = _lib.lib()
        dpool = _f32c(dpool, "dpool")
        dmu = torch.empty(ctx.layout.num_nodes, ctx.P, dtype=torch.float32, device=dpool.device)
        with torch.cuda.device(dpool.device):
            _lib.check(L.s2v_pool_backward(C.byref(ctx.layout.c_struct()), ctx.P, int(bool(ctx.norm_agg)), _p(dpool),
                                           _p(dmu), _stream()), "s2v_pool_backward")
        _lib.launch_count += 1
        return None, None, dmu


def _mask_u8(mask: torch.Tensor, n: int) -> torch.Tensor:
    _require_cuda(mask, "mask")
    mask = mask.reshape(-1)
    if mask.numel() != n:
        raise ValueError("mask has %d entries, expected %d" % (mask.numel(), n))
    if mask.dtype == torch.bool:
        return mask.contiguous().view(torch.uint8)
    return (mask != 0).view(torch.uint8)


class _MaskedSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layout, logits, mask_u8):
        L = _lib.lib()
        logits = _f32c(logits, "logits")
        prob = torch.empty_like(logits)
        with torch.cuda.device(logits.device):
            _lib.check(L.s2v_masked_softmax_forward(C.byref(layout.c_struct()), _p(logits), _p(mask_u8), _p(prob),
                                                    _stream()), "s2v_masked_softmax_forward")
        _lib.launch_count += 1
        ctx.layout = layout
        ctx.save_for_backward(prob)
        return prob

    @staticmethod
    def backward(ctx, dprob):
        L = _lib.lib()
        (prob,) = ctx.saved_tensors
        dprob = _f32c(dprob, "dprob")
        dlogits = torch.empty_like(prob)
        with torch.cuda.device(prob.device):
            _lib.check(L.s2v_masked_softmax_backward(C.byref(ctx.layout.c_struct()), _p(prob), _p(dprob), _p(dlogits),
                                                     _stream()), "s2v_masked_softmax_backward")
        _lib.launch_count += 1
        return None, dlogits, None


def _masked_argmax_raw(layout, q, mask_u8):
    L = _lib.lib()
    B = layout.num_graphs
    arg = torch.empty(B, dtype=torch.int64, device=q.device)
    val = torch.empty(B, dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(L.s2v_masked_argmax(C.byref(layout.c_struct()), _p(q), _p(mask_u8), _p(arg), _p(val), _stream()),
                   "s2v_masked_argmax")
    _lib.launch_count += 1
    return arg, val


class _MaskedMaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layout, q, mask_u8):
        q = _f32c(q, "q")
        arg, val = _masked_argmax_raw(layout, q, mask_u8)
        ctx.layout = layout
        ctx.save_for_backward(arg)
        ctx.mark_non_differentiable(arg)
        return val, arg

    @staticmethod
    def backward(ctx, dval, _darg):
        L = _lib.lib()
        (arg,) = ctx.saved_tensors
        dval = _f32c(dval, "dval")
        dq = torch.empty(ctx.layout.num_nodes, dtype=torch.float32, device=dval.device)
        with torch.cuda.device(dval.device):
            _lib.check(L.s2v_scatter_rows(C.byref(ctx.layout.c_struct()), _p(arg), _p(dval), _p(dq), _stream()),
                       "s2v_scatter_rows")
        _lib.launch_count += 1
        return None, dq, None


class _HeadSoftmaxFn(torch.autograd.Function):
    """logits = head(mu); prob = softmax(logits | mask) -- one launch forward when the fused kernel applies
    (models_basic_lstm.py:530-540).  Backward: masked-softmax backward, then the head backward."""

    @staticmethod
    def forward(ctx, layout, norm_agg, mask_u8, mu, *w):
        w = [_f32c(t.detach(), "parameter") for t in w]
        mu = _f32c(mu, "mu")
        if head_fused_eligible(layout, w):
            pool, gstate, _, prob = _head_fused_forward(layout, norm_agg, mu, w, "softmax", mask_u8)[:4]
        else:
            pool, gstate, q = _head_forward(layout, norm_agg, mu, w)
            prob = _MaskedSoftmaxFn.forward(_Ctx(), layout, q, mask_u8)
        ctx.layout, ctx.norm_agg = layout, norm_agg
        ctx.save_for_backward(mu, pool, gstate, prob, *w)
        return prob

    @staticmethod
    def backward(ctx, dprob):
        L = _lib.lib()
        mu, pool, gstate, prob, *w = ctx.saved_tensors
        dprob = _f32c(dprob, "dprob")
        dq = torch.empty_like(prob)
        with torch.cuda.device(prob.device):
            _lib.check(L.s2v_masked_softmax_backward(C.byref(ctx.layout.c_struct()), _p(prob), _p(dprob), _p(dq), _stream()),
                       "s2v_masked_softmax_backward")
        _lib.launch_count += 1
        dmu, grads, _ = _head_backward(ctx.layout, ctx.norm_agg, mu, w, pool, gstate, dq, False)
        return (None, None, None, dmu, *grads)


class _HeadMaxFn(torch.autograd.Function):
    """(max, argmax) of head(mu) over the reachable nodes of every graph (models_basic_lstm.py:548-564)."""

    @staticmethod
    def forward(ctx, layout, norm_agg, mask_u8, mu, *w):
        w = [_f32c(t.detach(), "parameter") for t in w]
        mu = _f32c(mu, "mu")
        if head_fused_eligible(layout, w):
            pool, gstate, _, _, arg, _, val = _head_fused_forward(layout, norm_agg, mu, w, "max", mask_u8)
        else:
            pool, gstate, q = _head_forward(layout, norm_agg, mu, w)
            arg, val = _masked_argmax_raw(layout, q, mask_u8)
        ctx.layout, ctx.norm_agg = layout, norm_agg
        ctx.save_for_backward(mu, pool, gstate, arg, *w)
        ctx.mark_non_differentiable(arg)
        return val, arg

    @staticmethod
    def backward(ctx, dval, _darg):
        L = _lib.lib()
        mu, pool, gstate, arg, *w = ctx.saved_tensors
        dval = _f32c(dval, "dval")
        dq = torch.empty(ctx.layout.num_nodes, dtype=torch.float32, device=dval.device)
        with torch.cuda.device(dval.device):
            _lib.check(L.s2v_scatter_rows(C.byref(ctx.layout.c_struct()), _p(arg), _p(dval), _p(dq), _stream()),
                       "s2v_scatter_rows")
        _lib.launch_count += 1
        dmu, grads, _ = _head_backward(ctx.layout, ctx.norm_agg, mu, w, pool, gstate, dq, False)
        return (None, None, None, dmu, *grads)


class _Ctx:
    """Stand-in context for calling an autograd.Function's forward as a plain function."""

    def save_for_backward(self, *a):
        pass


# ----------------------------------------------------------------------------------
# public functional API (the native contract of SURVEY.md §8b)
# ----------------------------------------------------------------------------------
def s2v_embed(layout: GraphLayout, x: torch.Tensor, embed_weights, T: int) -> torch.Tensor:
    """mu^T [num_nodes,p].  embed_weights: theta1a.w,b theta1b.w,b theta2.w,b theta3.w,b theta4.w,b."""
    return _EmbedFn.apply(layout, int(T), x, *embed_weights)         # x.requires_grad (lstm_type 'FE'): dx comes back too


def s2v_head(layout: GraphLayout, mu: torch.Tensor, head_weights, norm_agg: bool = True) -> torch.Tensor:
    """q / logits [num_nodes].  head_weights: theta5.w,b theta6.w,b theta7.w,b."""
    return _HeadFn.apply(layout, bool(norm_agg), mu, *head_weights)


def s2v_qnet(layout: GraphLayout, x: torch.Tensor, embed_weights, head_weights, T: int, norm_agg: bool = True):
    """Fused embedding + head: q [num_nodes] (QNet.forward, comb_opt.py:217-275)."""
    return _QNetFn.apply(layout, int(T), bool(norm_agg), x, *embed_weights, *head_weights)


def segment_pool(layout: GraphLayout, mu: torch.Tensor, norm_agg: bool = True) -> torch.Tensor:
    return _PoolFn.apply(layout, bool(norm_agg), mu)


def masked_softmax(layout: GraphLayout, logits: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Per-graph softmax with logits[~mask] = -inf (models_basic_lstm.py:539-540)."""
    flat = logits.reshape(-1)
    return _MaskedSoftmaxFn.apply(layout, flat, _mask_u8(mask, flat.numel())).view(logits.shape)


def masked_max(layout: GraphLayout, q: torch.Tensor, mask: torch.Tensor):
    """Per-graph (max value, local argmax node) over mask (models_basic_lstm.py:563-564)."""
    flat = q.reshape(-1)
    return _MaskedMaxFn.apply(layout, flat, _mask_u8(mask, flat.numel()))


@torch.no_grad()
def masked_argmax(layout: GraphLayout, q: torch.Tensor, mask: torch.Tensor):
    flat = _f32c(q.reshape(-1), "q")
    return _masked_argmax_raw(layout, flat, _mask_u8(mask, flat.numel()))


@torch.no_grad()
def list_argmax(layout: GraphLayout, q: torch.Tensor, reach_ptr: torch.Tensor, reach_idx: torch.Tensor):
    """Greedy action per graph over neighbour lists (comb_opt.py:319-326).
    reach_ptr int32 [B+1], reach_idx int32 local node ids.  Returns (pos, node, value)."""
    L = _lib.lib()
    q = _f32c(q.reshape(-1), "q")
    _require_cuda(reach_ptr, "reach_ptr")
    reach_ptr = reach_ptr.to(torch.int32).contiguous()
    reach_idx = reach_idx.to(device=q.device, dtype=torch.int32).contiguous()
    B = layout.num_graphs
    pos = torch.empty(B, dtype=torch.int64, device=q.device)
    node = torch.empty(B, dtype=torch.int64, device=q.device)
    val = torch.empty(B, dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(L.s2v_list_argmax(C.byref(layout.c_struct()), _p(q), _p(reach_ptr), _p(reach_idx), _p(pos), _p(node),
                                     _p(val), _stream()), "s2v_list_argmax")
    _lib.launch_count += 1
    return pos, node, val


def s2v_head_softmax(layout: GraphLayout, mu: torch.Tensor, head_weights, mask: torch.Tensor, norm_agg: bool = True):
    """prob [num_nodes] = softmax over the reachable nodes of every graph of the logit head (PPO pi,
    models_basic_lstm.py:530-540): head, invalid-action mask and the per-graph reduction in one launch where launches
    are the cost (csrc/s2v_head.cu:k_head_fused), head + mask kernels for large batches."""
    return _HeadSoftmaxFn.apply(layout, bool(norm_agg), _mask_u8(mask, layout.num_nodes), mu, *head_weights)


def s2v_head_max(layout: GraphLayout, mu: torch.Tensor, head_weights, mask: torch.Tensor, norm_agg: bool = True):
    """(max value [B], local argmax node [B]) of the Q head over the reachable nodes (PPO critic 'q',
    models_basic_lstm.py:548-564); differentiable in the value."""
    return _HeadMaxFn.apply(layout, bool(norm_agg), _mask_u8(mask, layout.num_nodes), mu, *head_weights)


@torch.no_grad()
def s2v_head_argmax(layout: GraphLayout, mu: torch.Tensor, head_weights, mask: torch.Tensor, norm_agg: bool = True):
    """Greedy evaluation (rl_policy.py:533-536): (local argmax node [B], value [B]); softmax is monotone, so the
    arg-max of the logits is the arg-max of the probabilities."""
    w = [_f32c(t.detach(), "parameter") for t in head_weights]
    mu = _f32c(mu, "mu")
    m8 = _mask_u8(mask, layout.num_nodes)
    if head_fused_eligible(layout, w):
        _, _, _, _, arg, _, val = _head_fused_forward(layout, bool(norm_agg), mu, w, "max", m8)
        return arg, val
    _, _, q = _head_forward(layout, bool(norm_agg), mu, w)
    return _masked_argmax_raw(layout, q, m8)


@torch.no_grad()
def s2v_head_list_argmax(layout: GraphLayout, mu: torch.Tensor, head_weights, reach_ptr: torch.Tensor,
                         reach_idx: torch.Tensor, norm_agg: bool = True):
    """DQN greedy action (comb_opt.py:319-326) from the embedding: (position in the neighbour list, node id, value,
    q) -- Q head and list arg-max in one launch when the fused kernel applies."""
    w = [_f32c(t.detach(), "parameter") for t in head_weights]
    mu = _f32c(mu, "mu")
    _require_cuda(reach_ptr, "reach_ptr")
    reach_ptr = reach_ptr.to(torch.int32).contiguous()
    reach_idx = reach_idx.to(device=mu.device, dtype=torch.int32).contiguous()
    if head_fused_eligible(layout, w):
        _, _, q, _, pos, node, val = _head_fused_forward(layout, bool(norm_agg), mu, w, "list", None, reach_ptr, reach_idx)
        return pos, node, val, q
    _, _, q = _head_forward(layout, bool(norm_agg), mu, w)
    pos, node, val = list_argmax(layout, q, reach_ptr, reach_idx)
    return pos, node, val, q


@torch.no_grad()
def s2v_qnet_list_argmax(layout: GraphLayout, x: torch.Tensor, embed_weights, head_weights, T: int, reach_ptr, reach_idx,
                         norm_agg: bool = True):
    """Acting (comb_opt.py:297-331): node features -> greedy action per graph over its neighbour list, in ONE launch
    for small problems (csrc/s2v_small.cu).  Returns (position in the list, node id, value) [B] each and q."""
    ew = [_f32c(t.detach(), "parameter") for t in embed_weights]
    hw = [_f32c(t.detach(), "parameter") for t in head_weights]
    x = _f32c(x, "x")
    _require_cuda(reach_ptr, "reach_ptr")
    reach_ptr = reach_ptr.to(torch.int32).contiguous()
    reach_idx = reach_idx.to(device=x.device, dtype=torch.int32).contiguous()
    if qnet_small_eligible(layout, ew):
        out = _qnet_small_forward(layout, int(T), x, ew, hw, bool(norm_agg), "list", None, reach_ptr, reach_idx)
        return out[6], out[7], out[8], out[4]
    _, mus = _embed_forward(layout, int(T), x, ew)
    return s2v_head_list_argmax(layout, mus[int(T) - 1], hw, reach_ptr, reach_idx, norm_agg)


@torch.no_grad()
def s2v_qnet_forward(layout: GraphLayout, x: torch.Tensor, embed_weights, head_weights, T: int, norm_agg: bool = True):
    """q [num_nodes] without autograd (QFunction.predict, comb_opt.py:297-307): one launch for small problems."""
    ew = [_f32c(t.detach(), "parameter") for t in embed_weights]
    hw = [_f32c(t.detach(), "parameter") for t in head_weights]
    x = _f32c(x, "x")
    if qnet_small_eligible(layout, ew):
        return _qnet_small_forward(layout, int(T), x, ew, hw, bool(norm_agg), "q")[4]
    _, mus = _embed_forward(layout, int(T), x, ew)
    if head_fused_eligible(layout, hw):
        return _head_fused_forward(layout, bool(norm_agg), mus[int(T) - 1], hw, "q")[2]
    return _head_forward(layout, bool(norm_agg), mus[int(T) - 1], hw)[2]
