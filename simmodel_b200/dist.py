"""Data parallelism over independent graphs.

Graphs never interact (block-diagonal batch, per-graph pooling and masks), so a batch is
split contiguously by graph across ranks with no data-path collective.  Parameters are
replicated; the only exchange is ONE all-reduce of a flat fp32 gradient buffer per optimiser
step (86 KB for the s2v Q-net), NCCL over NVLink on GPUs, gloo in the CPU tests.

Loss normalisation follows the reference's mean over the *global* batch (SmoothL1 mean over
B, comb_opt.py:359; loss_tsr.mean() over all steps of all rollouts, models_basic_lstm.py:650):
each rank's gradient of its local mean is weighted by local_count / global_count and summed.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_graphs(num_graphs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced split of graph ids [lo, hi) (first `rem` ranks get one more)."""
    q, rem = divmod(num_graphs, world_size)
    lo = rank * q + min(rank, rem)
    return lo, lo + q + (1 if rank < rem else 0)


def shard_graphs_by_nodes(sizes: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous split balancing the number of NODES per rank (variable-size graphs):
    greedy cut at the running-sum targets k * total / world."""
    total = float(sum(sizes))
    bounds, acc, g = [0], 0.0, 0
    for k in range(1, world_size):
        target = total * k / world_size
        while g < len(sizes) and acc + sizes[g] / 2.0 <= target:
            acc += sizes[g]
            g += 1
        bounds.append(g)
    bounds.append(len(sizes))
    return [(bounds[i], bounds[i + 1]) for i in range(world_size)]


def flatten_grads(params: Iterable[torch.nn.Parameter]) -> Tuple[torch.Tensor, List[torch.nn.Parameter]]:
    ps = [p for p in params if p.requires_grad]
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps])
    return flat, ps


def unflatten_grads(flat: torch.Tensor, ps: List[torch.nn.Parameter]) -> None:
    off = 0
    for p in ps:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n


def shared_flat_grad(ps: List[torch.nn.Parameter]):
    """The gradients of ``ps`` as ONE tensor when they already are views that tile a single buffer back to back in
    this order (the kernels write a flat [embed | head] gradient and autograd keeps the views as ``.grad``,
    ops._QNetFn.backward); None otherwise."""
    if not ps or any(p.grad is None or not p.grad.is_contiguous() or p.grad.dtype != ps[0].grad.dtype for p in ps):
        return None
    g0 = ps[0].grad
    st = g0.untyped_storage()
    esz = g0.element_size()
    off = g0.storage_offset()
    for p in ps:
        g = p.grad
        if g.untyped_storage().data_ptr() != st.data_ptr() or g.storage_offset() != off:
            return None
        off += g.numel()
    n = off - g0.storage_offset()
    if (g0.storage_offset() + n) * esz > st.nbytes():
        return None
    return torch.empty(0, dtype=g0.dtype, device=g0.device).set_(st, g0.storage_offset(), (n,), (1,))


def allreduce_gradients(params: Iterable[torch.nn.Parameter], local_count: int, global_count: int, group=None) -> torch.Tensor:
    """One all-reduce of the local_count/global_count-weighted gradients; returns the flat buffer.
    When the gradients already live in one flat buffer (the s2v Q-net: 86 KB) it is reduced IN PLACE -- no cat, no
    copies back: one scale + one NCCL call (or a single ReduceOp.AVG call when every rank holds the same count)."""
    ps = [p for p in params if p.requires_grad]
    active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    flat = shared_flat_grad(ps)
    if flat is not None:
        if active:
            world = dist.get_world_size(group)
            if local_count * world == global_count and dist.get_backend(group) == "nccl":
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
            else:
                flat.mul_(float(local_count) / float(global_count))
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return flat
    flat, ps = flatten_grads(ps)
    if active:
        flat.mul_(float(local_count) / float(global_count))
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        unflatten_grads(flat, ps)
    return flat


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)
