"""`torch.library` registration of the C-ABI kernels: namespace ``s2v_b200``.

The `autograd.Function`s of ``ops.py`` are what the nn.Module mirrors call; this module exposes the same entry points
as dispatcher-level custom ops (schema, CUDA implementation, fake/meta implementation, autograd formula), which makes
them visible to ``torch.compile`` / ``torch.export`` / ``opcheck`` and callable from C++ or TorchScript-free serving
stacks through ``torch.ops.s2v_b200.*``.  Arguments are plain tensors and ints -- the graph layout travels as its
int32 arrays ("thin C-ABI torch custom-op layer"); nothing here computes, every op forwards to the
ctypes binding of ``include/s2v_b200.h``.  CUDA tensors only: there is no CPU implementation to dispatch to.

    rowptr, col, rowptr_t, col_t, indeg = torch.ops.s2v_b200.csr_build(edge_index, n)
    q = torch.ops.s2v_b200.qnet(x, weights16, T, norm_agg, rowptr, col, rowptr_t, col_t, indeg, ptr, gid, npad, B, period, npad_scalar)
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import ops
from .layout import GraphLayout

NS = "s2v_b200"


def _layout(rowptr, col, rowptr_t, col_t, indeg, ptr, gid, npad, num_nodes, num_graphs, period, npad_scalar) -> GraphLayout:
    return GraphLayout(num_nodes, num_graphs, period, rowptr, col, rowptr_t, col_t, indeg, ptr, gid,
                       npad if npad is not None else int(npad_scalar))


def layout_args(layout: GraphLayout):
    """GraphLayout -> the (tensor..., int...) tail every op of this namespace takes."""
    npad_t = layout.npad if isinstance(layout.npad, torch.Tensor) else None
    npad_s = 0 if npad_t is not None else int(layout.npad)
    return (layout.rowptr, layout.col, layout.rowptr_t, layout.col_t, layout.indeg, layout.ptr, layout.gid, npad_t,
            layout.num_nodes, layout.num_graphs, layout.period, npad_s)


# ---------------------------------------------------------------------------------------------------------------
# K1: CSR builder
# ---------------------------------------------------------------------------------------------------------------
@torch.library.custom_op(NS + "::csr_build", mutates_args=())
def csr_build(edge_index: Tensor, n: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    return GraphLayout.build_csr(edge_index, int(n), validate=False)


@csr_build.register_fake
def _(edge_index, n):
    E = edge_index.shape[1]
    i32 = dict(dtype=torch.int32, device=edge_index.device)
    return (torch.empty(n + 1, **i32), torch.empty(E, **i32), torch.empty(n + 1, **i32), torch.empty(E, **i32),
            torch.empty(n, **i32))


# ---------------------------------------------------------------------------------------------------------------
# K2-K5: Q-network forward / backward (embedding + head), the op QNet.forward_graph is made of
# ---------------------------------------------------------------------------------------------------------------
@torch.library.custom_op(NS + "::qnet_fwd", mutates_args=())
def qnet_fwd(x: Tensor, weights: List[Tensor], T: int, norm_agg: bool, rowptr: Tensor, col: Tensor, rowptr_t: Tensor,
             col_t: Tensor, indeg: Tensor, ptr: Optional[Tensor], gid: Optional[Tensor], npad: Optional[Tensor],
             num_nodes: int, num_graphs: int, period: int, npad_scalar: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> (q [n], base [n,p], mus [T,n,p], pool [B,p], gstate [B,p]); the last four feed qnet_bwd."""
    lay = _layout(rowptr, col, rowptr_t, col_t, indeg, ptr, gid, npad, num_nodes, num_graphs, period, npad_scalar)
    w = [ops._f32c(t, "parameter") for t in weights]
    x = ops._f32c(x, "x")
    if ops.qnet_small_eligible(lay, w[:10]):
        base, mus, pool, gstate, q = ops._qnet_small_forward(lay, int(T), x, w[:10], w[10:], bool(norm_agg), "q")[:5]
    else:
        base, mus = ops._embed_forward(lay, int(T), x, w[:10])
        if ops.head_fused_eligible(lay, w[10:]):
            pool, gstate, q = ops._head_fused_forward(lay, bool(norm_agg), mus[int(T) - 1], w[10:], "q")[:3]
        else:
            pool, gstate, q = ops._head_forward(lay, bool(norm_agg), mus[int(T) - 1], w[10:])
    return q, base, mus, pool, gstate


@qnet_fwd.register_fake
def _(x, weights, T, norm_agg, rowptr, col, rowptr_t, col_t, indeg, ptr, gid, npad, num_nodes, num_graphs, period, npad_scalar):
    p = weights[0].shape[0]
    f = dict(dtype=torch.float32, device=x.device)
    return (torch.empty(num_nodes, **f), torch.empty(num_nodes, p, **f), torch.empty(T, num_nodes, p, **f),
            torch.empty(num_graphs, p, **f), torch.empty(num_graphs, p, **f))


@torch.library.custom_op(NS + "::qnet_bwd", mutates_args=())
def qnet_bwd(dq: Tensor, x: Tensor, weights: List[Tensor], base: Tensor, mus: Tensor, pool: Tensor, gstate: Tensor, T: int,
             norm_agg: bool, rowptr: Tensor, col: Tensor, rowptr_t: Tensor, col_t: Tensor, indeg: Tensor,
             ptr: Optional[Tensor], gid: Optional[Tensor], npad: Optional[Tensor], num_nodes: int, num_graphs: int,
             period: int, npad_scalar: int) -> Tensor:
    """-> flat gradient [embed | head] in QNet.parameters() order (s2v_embed_grad_floats + s2v_head_grad_floats)."""
    from . import _lib
    lay = _layout(rowptr, col, rowptr_t, col_t, indeg, ptr, gid, npad, num_nodes, num_graphs, period, npad_scalar)
    w = [ops._f32c(t, "parameter") for t in weights]
    L = _lib.lib()
    T = int(T)
    P, F = int(w[0].shape[0]), int(w[0].shape[1])
    ne, nh = int(L.s2v_embed_grad_floats(F, P)), int(L.s2v_head_grad_floats(P))
    flat = torch.empty(ne + nh, dtype=torch.float32, device=x.device)
    dus = torch.empty(T, num_nodes, P, dtype=torch.float32, device=x.device)
    ops._head_backward(lay, bool(norm_agg), mus[T - 1], w[10:], pool, gstate, ops._f32c(dq, "dq"), True, dmu_out=dus[T - 1],
                       grad_out=flat[ne:])
    ops._embed_backward(lay, T, x, w[:10], base, mus, dus[T - 1], True, dus=dus, grad_out=flat[:ne])
    return flat


@qnet_bwd.register_fake
def _(dq, x, weights, base, mus, pool, gstate, T, norm_agg, rowptr, col, rowptr_t, col_t, indeg, ptr, gid, npad, num_nodes,
      num_graphs, period, npad_scalar):
    return torch.empty(sum(int(w.numel()) for w in weights), dtype=torch.float32, device=x.device)


def _qnet_setup(ctx, inputs, output):
    x, weights, T, norm_agg, *g = inputs
    _, base, mus, pool, gstate = output
    ctx.T, ctx.norm_agg, ctx.g_ints = T, norm_agg, g[8:]
    ctx.nw = len(weights)
    ctx.shapes = [tuple(w.shape) for w in weights]
    ctx.save_for_backward(x, base, mus, pool, gstate, *weights, *[t for t in g[:8] if t is not None])
    ctx.g_present = [t is not None for t in g[:8]]


def _qnet_backward(ctx, dq, *_unused):
    x, base, mus, pool, gstate, *rest = ctx.saved_tensors
    weights, gt = rest[:ctx.nw], list(rest[ctx.nw:])
    g = [gt.pop(0) if present else None for present in ctx.g_present]
    flat = torch.ops.s2v_b200.qnet_bwd(dq.contiguous(), x, list(weights), base, mus, pool, gstate, ctx.T, ctx.norm_agg, *g, *ctx.g_ints)
    grads, off = [], 0
    for shp in ctx.shapes:
        n = 1
        for d in shp:
            n *= d
        grads.append(flat[off:off + n].view(shp))
        off += n
    return (None, grads, None, None) + (None,) * 12


qnet_fwd.register_autograd(_qnet_backward, setup_context=_qnet_setup)


def qnet(x: Tensor, weights: List[Tensor], T: int, norm_agg: bool, layout: GraphLayout) -> Tensor:
    """q [sum N] through the dispatcher (same kernels as QNet.forward_graph; differentiable in the 16 weights)."""
    return torch.ops.s2v_b200.qnet_fwd(x, list(weights), int(T), bool(norm_agg), *layout_args(layout))[0]


# ---------------------------------------------------------------------------------------------------------------
# K6/K7: invalid-action mask + per-graph reductions
# ---------------------------------------------------------------------------------------------------------------
class _RangeLayout:
    """Graph ranges only (no CSR): what the mask / reduction kernels read of s2v_graph_t."""

    def __init__(self, ptr, num_nodes, num_graphs, period):
        import ctypes as C
        from . import _lib
        self.num_nodes, self.num_graphs, self.period = int(num_nodes), int(num_graphs), int(period)
        self.ptr = None if period > 0 else ptr.to(torch.int32).contiguous()
        g = _lib.Graph()
        g.num_nodes, g.num_graphs, g.period, g.num_edges = self.num_nodes, self.num_graphs, self.period, 0
        g.ptr = C.c_void_p(0 if self.ptr is None else self.ptr.data_ptr())
        self._g = g

    def c_struct(self):
        return self._g


def _ranges(ptr, num_nodes, num_graphs, period):
    return _RangeLayout(ptr, num_nodes, num_graphs, period)


@torch.library.custom_op(NS + "::masked_softmax", mutates_args=())
def masked_softmax(logits: Tensor, mask: Tensor, ptr: Tensor, num_graphs: int, period: int) -> Tensor:
    """Per-graph softmax with logits[~mask] = -inf (models_basic_lstm.py:539-540).  ptr int32 [B+1] (ignored when
    period > 0: every graph has `period` nodes)."""
    lay = _ranges(ptr, int(logits.numel()), int(num_graphs), int(period))
    flat = ops._f32c(logits.reshape(-1), "logits")
    return ops._MaskedSoftmaxFn.forward(ops._Ctx(), lay, flat, ops._mask_u8(mask, flat.numel())).view(logits.shape)


@masked_softmax.register_fake
def _(logits, mask, ptr, num_graphs, period):
    return torch.empty_like(logits)


@torch.library.custom_op(NS + "::masked_softmax_bwd", mutates_args=())
def masked_softmax_bwd(prob: Tensor, dprob: Tensor, ptr: Tensor, num_graphs: int, period: int) -> Tensor:
    import ctypes as C
    from . import _lib
    from .layout import _p, _stream
    lay = _ranges(ptr, int(prob.numel()), int(num_graphs), int(period))
    p_, d_ = ops._f32c(prob.reshape(-1), "prob"), ops._f32c(dprob.reshape(-1), "dprob")
    out = torch.empty_like(p_)
    with torch.cuda.device(p_.device):
        _lib.check(_lib.lib().s2v_masked_softmax_backward(C.byref(lay.c_struct()), _p(p_), _p(d_), _p(out), _stream()),
                   "s2v_masked_softmax_backward")
    return out.view(prob.shape)


@masked_softmax_bwd.register_fake
def _(prob, dprob, ptr, num_graphs, period):
    return torch.empty_like(prob)


def _ms_setup(ctx, inputs, output):
    _, _, ptr, num_graphs, period = inputs
    ctx.num_graphs, ctx.period = num_graphs, period
    ctx.save_for_backward(output, ptr)


def _ms_backward(ctx, dprob):
    prob, ptr = ctx.saved_tensors
    return torch.ops.s2v_b200.masked_softmax_bwd(prob, dprob.contiguous(), ptr, ctx.num_graphs, ctx.period), None, None, None, None


masked_softmax.register_autograd(_ms_backward, setup_context=_ms_setup)


@torch.library.custom_op(NS + "::masked_argmax", mutates_args=())
def masked_argmax(q: Tensor, mask: Tensor, ptr: Tensor, num_graphs: int, period: int) -> Tuple[Tensor, Tensor]:
    """(local arg-max node int64 [B], value [B]) over mask != 0; first maximum wins (models_basic_lstm.py:563-564)."""
    lay = _ranges(ptr, int(q.numel()), int(num_graphs), int(period))
    flat = ops._f32c(q.reshape(-1), "q")
    return ops._masked_argmax_raw(lay, flat, ops._mask_u8(mask, flat.numel()))


@masked_argmax.register_fake
def _(q, mask, ptr, num_graphs, period):
    return (torch.empty(num_graphs, dtype=torch.int64, device=q.device), torch.empty(num_graphs, dtype=torch.float32, device=q.device))


OPS = ("csr_build", "qnet_fwd", "qnet_bwd", "masked_softmax", "masked_softmax_bwd", "masked_argmax")
