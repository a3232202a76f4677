"""Second, independently written restatement of the GATv2 layer (TEST INFRASTRUCTURE ONLY).

oracle/gat_ref.py restates PyG's GATv2Conv the way PyG computes it: an edge list with self loops appended, per-edge
scores, a segment softmax and a scatter-add.  The kernels are checked against it -- and it cannot be checked against
PyG itself, which is neither under /root/reference nor installable in this image.  A shared misreading (self-loop
handling, which of the two projections is attended over, bias placement, the softmax's denominator) would go unnoticed.
This file derives the same layer from the PAPER's formula instead (Brody, Alon, Yahav: "How Attentive are Graph
Attention Networks?", ICLR 2022, eq. 7, with the two-matrix parametrisation W = [W_l | W_r] of its reference
implementation), using none of the machinery above: a dense N x N neighbourhood matrix, a dense masked softmax, a
dense matrix product.

    e_ij   = a_h . LeakyReLU_0.2( W_r x_i + b_r  +  W_l x_j + b_l )            i = target, j = source, per head h
    N(i)   = { j : (j -> i) is an edge, j != i }  U  { i }                     every node attends to itself exactly once
    alpha  = softmax over j in N(i) of e_ij                                    (multi-edges count with their multiplicity)
    out_i  = concat_h  sum_j alpha_ij (W_l x_j + b_l)_h   +  bias

tests/test_gat_oracle.py asserts that the two restatements agree to fp64 rounding on every GAT golden case (which
also shows that PyG's `+ 1e-16` in the softmax denominator, kept in gat_ref.py and in the kernels, is immaterial)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def neighbourhood_matrix(edge_index: torch.Tensor, n: int, dtype=torch.float64) -> torch.Tensor:
    """M[i, j] = number of edges j -> i with j != i, and M[i, i] = 1."""
    M = torch.zeros(n, n, dtype=dtype)
    src, dst = edge_index[0], edge_index[1]
    off = src != dst
    M.index_put_((dst[off], src[off]), torch.ones(int(off.sum()), dtype=dtype), accumulate=True)
    M += torch.eye(n, dtype=dtype)
    return M


def gatv2_layer_dense(x, M, W_l, b_l, W_r, b_r, att, bias, rows_per_chunk=256):
    """x [N, in]; M [N, N] neighbourhood multiplicities; att [1, H, C] -> [N, H*C]."""
    N = x.shape[0]
    H, C = att.shape[1], att.shape[2]
    zl = (x @ W_l.t() + b_l).view(N, H, C)          # what a source contributes (and what is aggregated)
    zr = (x @ W_r.t() + b_r).view(N, H, C)          # what the target contributes to the score
    out = torch.empty(N, H, C, dtype=x.dtype)
    a = att.view(1, 1, H, C)
    for i0 in range(0, N, rows_per_chunk):
        i1 = min(N, i0 + rows_per_chunk)
        s = F.leaky_relu(zr[i0:i1, None, :, :] + zl[None, :, :, :], 0.2)        # [rows, N, H, C]
        e = (s * a).sum(-1)                                                       # [rows, N, H]
        m = M[i0:i1, :, None]
        e = e.masked_fill(m == 0, float("-inf"))
        e = e - e.max(dim=1, keepdim=True).values
        w = m * e.exp()                                                           # multiplicity-weighted
        alpha = w / w.sum(dim=1, keepdim=True)
        out[i0:i1] = torch.einsum("ijh,jhc->ihc", alpha, zl)
    return out.reshape(N, H * C) + bias


def gat_forward_dense(P, x, edge_index, num_layers, prefix="gat."):
    """The stack as the reference builds it (BasicGNN: ReLU between layers, none after the last)."""
    M = neighbourhood_matrix(edge_index, x.shape[0], x.dtype)
    for k in range(num_layers):
        g = lambda s: P["%sconvs.%d.%s" % (prefix, k, s)]
        x = gatv2_layer_dense(x, M, g("lin_l.weight"), g("lin_l.bias"), g("lin_r.weight"), g("lin_r.bias"), g("att"), g("bias"))
        if k < num_layers - 1:
            x = torch.relu(x)
    return x
