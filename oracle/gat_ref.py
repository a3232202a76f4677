"""CPU restatement of the GATv2 embedding branch (TEST INFRASTRUCTURE ONLY -- never imported by simmodel_b200/).

SURVEY.md §8f rank 1: ``--qnet gat2``.  The reference builds the embedding from two THIRD-PARTY classes that are
absent from /root/reference and from this image:

  * torch_geometric.nn.conv.GATv2Conv and torch_geometric.nn.models.basic_gnn.BasicGNN (unpinned ``pip install
    torch-geometric``, README.md:22, python 3.8 / torch 1.11 era => PyG ~2.0.4).  Call sites: the ``GATv2(BasicGNN)``
    subclasses at modules/gnn/comb_opt.py:48-62 and modules/ppo/models_basic_lstm.py:326-338, instantiated at
    comb_opt.py:163-171 (DQN: hidden = emb_dim, heads 2, 5 layers) and models_basic_lstm.py:414-426 (PPO:
    hidden = 2*emb_dim, heads = config['gat_heads'] = 4, emb_iterT layers, out = emb_dim; ppo_custom.py:255-257),
    called as ``self.gat(pyg_data.x, pyg_data.edge_index)`` (comb_opt.py:94; models_basic_lstm.py:513).

**Parity status: the conv arithmetic is UNPINNED against real PyG** (it cannot be installed here; no golden vectors
exist in the reference).  What is pinned: (i) parameter names and shapes, against the reference's shipped gat2
checkpoint (results/.../SEED2205/checkpoints/best_model.tar -> tests/golden/gat2_ppo_checkpoint.npz); (ii) the
reference's own code around the conv -- ``QNet_GAT.forward`` (comb_opt.py:84-135), ``create_embeddings / pi / v``
(models_basic_lstm.py:507-566) -- which oracle/gen_golden_gat.py runs UNMODIFIED with this restatement injected in
place of the absent PyG classes.

Published algorithm restated (PyG 2.0.4, torch_geometric/nn/conv/gatv2_conv.py, utils/softmax.py, models/basic_gnn.py):

  x_l = lin_l(x).view(N,H,C);  x_r = lin_r(x).view(N,H,C)          (share_weights=False: two Linear layers with bias)
  edge_index <- remove_self_loops(edge_index); add_self_loops(num_nodes=N)   (self loops appended after the other edges)
  per edge (j -> i) = (edge_index[0], edge_index[1]):  e = sum_c att[h,c] * leaky_relu(x_r[i] + x_l[j], 0.2)[h,c]
  alpha = exp(e - max_i) / (sum_i exp(e - max_i) + 1e-16)         (softmax over the incoming edges of target i, per head)
  out[i] = sum_j alpha * x_l[j]   (scatter-add in edge order);  concat heads -> view(N, H*C);  + bias
  BasicGNN: x = conv_k(x); ReLU (dropout p=0) between layers, nothing after the last (jk=None, norm=None).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def remove_then_add_self_loops(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """PyG remove_self_loops + add_self_loops: surviving edges keep their order, loops 0..N-1 are appended."""
    keep = edge_index[0] != edge_index[1]
    loops = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat((edge_index[:, keep], torch.stack((loops, loops))), dim=1)


def segment_softmax(src: torch.Tensor, index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """torch_geometric.utils.softmax (2.0.4): max-shifted, denominator + 1e-16.  src [E,H], index [E]."""
    H = src.shape[1]
    idx = index.view(-1, 1).expand(-1, H)
    src_max = torch.full((num_nodes, H), float("-inf"), dtype=src.dtype).scatter_reduce(0, idx, src, "amax", include_self=True)
    out = (src - src_max.detach()[index]).exp()
    out_sum = torch.zeros(num_nodes, H, dtype=src.dtype).index_add_(0, index, out)
    return out / (out_sum[index] + 1e-16)


def gatv2_conv(x, edge_index, lin_l_w, lin_l_b, lin_r_w, lin_r_b, att, bias, negative_slope=0.2):
    """One GATv2Conv layer (concat=True, add_self_loops=True, share_weights=False, dropout 0).
    x [N,in]; lin_*_w [H*C,in]; att [1,H,C]; bias [H*C] -> [N,H*C]."""
    N = x.shape[0]
    H, C = att.shape[1], att.shape[2]
    x_l = F.linear(x, lin_l_w, lin_l_b).view(N, H, C)
    x_r = F.linear(x, lin_r_w, lin_r_b).view(N, H, C)
    ei = remove_then_add_self_loops(edge_index, N)
    src, dst = ei[0], ei[1]
    s = F.leaky_relu(x_r[dst] + x_l[src], negative_slope)
    e = (s * att).sum(dim=-1)                              # [E,H]
    alpha = segment_softmax(e, dst, N)
    msg = x_l[src] * alpha.unsqueeze(-1)
    out = torch.zeros(N, H, C, dtype=x.dtype).index_add_(0, dst, msg)
    return out.view(N, H * C) + bias


def gat_layer_dims(in_channels, hidden_channels, num_layers, out_channels, heads):
    """(in, out_total, heads_k) per layer as the reference's GATv2.init_conv resolves them
    (comb_opt.py:52-62: heads falls back to 1 when it does not divide the layer width; concat=True)."""
    widths = [hidden_channels] * (num_layers - 1) + [out_channels]
    dims, cin = [], in_channels
    for w in widths:
        h = heads if w % heads == 0 else 1
        dims.append((cin, w, h))
        cin = w
    return dims


def conv_param_names(k):
    return ["gat.convs.%d.%s" % (k, s) for s in ("att", "bias", "lin_l.weight", "lin_l.bias", "lin_r.weight", "lin_r.bias")]


def gat_forward(P, x, edge_index, num_layers, prefix="gat."):
    """BasicGNN.forward with GATv2Conv layers.  P: state-dict style {name: tensor}."""
    for k in range(num_layers):
        g = lambda s: P["%sconvs.%d.%s" % (prefix, k, s)]
        x = gatv2_conv(x, edge_index, g("lin_l.weight"), g("lin_l.bias"), g("lin_r.weight"), g("lin_r.bias"), g("att"), g("bias"))
        if k < num_layers - 1:
            x = F.relu(x)
    return x


def head(P, mu, ptr, norm_agg=True, suffix=""):
    """comb_opt.py:100-125 (= the s2v head, 261-273) on the GAT embedding."""
    sizes = (ptr[1:] - ptr[:-1])
    B = sizes.numel()
    batch = torch.repeat_interleave(torch.arange(B), sizes)
    pool = torch.zeros(B, mu.shape[1], dtype=mu.dtype).index_add_(0, batch, mu)
    if norm_agg:
        pool = pool / sizes.clamp(min=1).to(mu.dtype).unsqueeze(1)
    gs = F.linear(pool[batch], P["theta6%s.weight" % suffix], P["theta6%s.bias" % suffix])   # expanded first (:108-113)
    la = F.linear(mu, P["theta7%s.weight" % suffix], P["theta7%s.bias" % suffix])
    rep = F.relu(torch.cat((gs, la), dim=1))
    return F.linear(rep, P["theta5%s.weight" % suffix], P["theta5%s.bias" % suffix]).squeeze(-1)


def qnet_gat_forward(P, x, edge_index, ptr, num_layers=5, norm_agg=True):
    """QNet_GATv2.forward (comb_opt.py:84-127): q over all nodes of the batch."""
    mu = gat_forward(P, x, edge_index, num_layers)
    return head(P, mu, ptr, norm_agg)


def ppo_heads(P, mu, reachable, critic):
    """pi / v on top of the embedding mu [S,N,p] (models_basic_lstm.py:530-540, 548-564): equal-size graphs,
    plain mean pooling, reachable-node mask, softmax over nodes / max over reachable nodes."""
    S, N, p = mu.shape
    ptr = torch.arange(S + 1, dtype=torch.int64) * N
    logits = head(P, mu.reshape(S * N, p), ptr, True, "_pi").view(S, N)
    logits = logits.masked_fill(~reachable, float("-inf"))
    prob = F.softmax(logits, dim=1)
    if critic == "v":
        v = F.linear(mu.mean(dim=1, keepdim=True), P["theta5_v.weight"], P["theta5_v.bias"]).squeeze(-1)
    else:
        qv = head(P, mu.reshape(S * N, p), ptr, True, "_v").view(S, N).masked_fill(~reachable, float("-inf"))
        v = qv.max(dim=1)[0].unsqueeze(-1)
    return prob, v


class Data:
    """torch_geometric.data.Data(x, edge_index) as the gat2 branch uses it (models_basic_lstm.py:511)."""

    def __init__(self, x=None, edge_index=None):
        self.x, self.edge_index = x, edge_index
        self.num_nodes = int(x.shape[0])


class _Batch:
    pass


def batch_from_data_list(data_list):
    """Batch.from_data_list: x concatenated, edge_index offset by the cumulative node counts, batch, ptr."""
    b = _Batch()
    sizes = torch.tensor([int(d.x.shape[0]) for d in data_list], dtype=torch.int64)
    b.ptr = torch.zeros(len(data_list) + 1, dtype=torch.int64)
    b.ptr[1:] = torch.cumsum(sizes, 0)
    b.batch = torch.repeat_interleave(torch.arange(len(data_list)), sizes)
    b.x = torch.cat([d.x for d in data_list], dim=0)
    b.edge_index = torch.cat([d.edge_index.to(torch.int64) + int(b.ptr[k]) for k, d in enumerate(data_list)], dim=1)
    b.num_graphs = len(data_list)
    b.num_nodes = int(b.ptr[-1])
    return b


# ---- module stand-ins injected for the absent PyG classes when the unmodified reference is imported ----------------
class PygLinear(nn.Module):
    """torch_geometric.nn.dense.linear.Linear: glorot weight, bias ~ U(-1/sqrt(in), 1/sqrt(in))."""

    def __init__(self, in_channels, out_channels, bias=True, weight_initializer="glorot"):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        a = math.sqrt(6.0 / (in_channels + out_channels))
        nn.init.uniform_(self.weight, -a, a)
        if bias:
            b = 1.0 / math.sqrt(in_channels)
            nn.init.uniform_(self.bias, -b, b)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)


class GATv2Conv(nn.Module):
    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0,
                 add_self_loops=True, bias=True, share_weights=False, **kwargs):
        super().__init__()
        assert concat and add_self_loops and bias and dropout == 0.0, "only the reference's configurations"
        self.heads, self.out_channels, self.negative_slope = heads, out_channels, negative_slope
        self.lin_l = PygLinear(in_channels, heads * out_channels)
        self.lin_r = self.lin_l if share_weights else PygLinear(in_channels, heads * out_channels)   # gatv2_conv.py: shared module
        self.att = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels))
        a = math.sqrt(6.0 / (heads + out_channels))
        nn.init.uniform_(self.att, -a, a)

    def forward(self, x, edge_index):
        return gatv2_conv(x, edge_index, self.lin_l.weight, self.lin_l.bias, self.lin_r.weight, self.lin_r.bias,
                          self.att, self.bias, self.negative_slope)


class MessagePassing(nn.Module):
    pass


class BasicGNN(nn.Module):
    """torch_geometric.nn.models.basic_gnn.BasicGNN restricted to what the reference uses
    (norm=None, jk=None, dropout 0, act ReLU)."""

    def __init__(self, in_channels, hidden_channels, num_layers, out_channels=None, dropout=0.0, **kwargs):
        super().__init__()
        self.in_channels, self.hidden_channels, self.num_layers = in_channels, hidden_channels, num_layers
        self.out_channels = out_channels if out_channels is not None else hidden_channels
        self.dropout = dropout
        self.convs = nn.ModuleList()
        self.convs.append(self.init_conv(in_channels, hidden_channels, **kwargs))
        for _ in range(num_layers - 2):
            self.convs.append(self.init_conv(hidden_channels, hidden_channels, **kwargs))
        self.convs.append(self.init_conv(hidden_channels, self.out_channels, **kwargs))

    def init_conv(self, in_channels, out_channels, **kwargs):
        raise NotImplementedError

    def forward(self, x, edge_index):
        for i in range(self.num_layers):
            x = self.convs[i](x, edge_index)
            if i == self.num_layers - 1:
                break
            x = F.relu(x)
        return x
