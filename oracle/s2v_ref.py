"""CPU restatement of the reference s2v hot path (TEST INFRASTRUCTURE ONLY).

Two restatements of the same arithmetic, both plain torch on CPU, both usable
in fp32 and fp64:

* ``*_dense``  : follows the reference line by line (dense padded adjacency,
  the (B,Npad,Npad,p) theta4 tensor, un-padding by a boolean list,
  scatter_mean pooling).  Ground truth for semantics.
    - DQN   : modules/gnn/comb_opt.py:217-275 (QNet.forward),
              309-331 (greedy action), 355-359 (loss gather)
    - PPO   : modules/ppo/models_basic_lstm.py:479-505 (s2v), 507-524
              (create_embeddings), 526-542 (pi), 544-566 (v)
* ``*_sparse`` : the (x, edge_index, ptr, n_pad) form the CUDA kernels
  implement (closed-form theta4 term from in-degrees, index_add aggregation).
  Scales to sizes where the dense form needs terabytes.

Third-party arithmetic the reference calls that is not under /root/reference
is restated here from its published semantics:
  torch_scatter (2.0.9 era)  scatter_add / scatter_mean (sum / count clamped >= 1)
  torch_geometric (~2.0.4)   Batch.from_data_list (ptr, batch), dense_to_sparse
                             (= W.nonzero().t()), to_dense_adj (scatter-add of ones)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

QNET_LINEARS = ("theta1a", "theta1b", "theta2", "theta3", "theta4", "theta5", "theta6", "theta7")


# ----------------------------------------------------------------------------
# third-party restatements
# ----------------------------------------------------------------------------
def scatter_add(src: torch.Tensor, index: torch.Tensor, dim_size: int | None = None) -> torch.Tensor:
    """torch_scatter.scatter_add(src, index, dim=0) (call site comb_opt.py:266)."""
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype)
    return out.index_add_(0, index, src)


def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim_size: int | None = None) -> torch.Tensor:
    """torch_scatter.scatter_mean(src, index, dim=0) (call site comb_opt.py:262):
    sum divided by the per-segment count clamped to >= 1."""
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    out = scatter_add(src, index, dim_size)
    cnt = torch.zeros(dim_size, dtype=src.dtype).index_add_(0, index, torch.ones(index.numel(), dtype=src.dtype))
    cnt = cnt.clamp(min=1)
    return out / cnt.view((-1,) + (1,) * (src.dim() - 1))


def to_dense_adj(edge_index: torch.Tensor, max_num_nodes: int | None = None) -> torch.Tensor:
    """torch_geometric.utils.to_dense_adj(edge_index) for one graph -> (1,N,N) float32.
    Duplicate edges add up (scatter-add of ones). Call site models_basic_lstm.py:515."""
    n = int(edge_index.max()) + 1 if edge_index.numel() else 0
    if max_num_nodes is not None:
        n = max_num_nodes
    adj = torch.zeros(n * n, dtype=torch.float32)
    adj.index_add_(0, edge_index[0] * n + edge_index[1], torch.ones(edge_index.shape[1]))
    return adj.view(1, n, n)


def dense_to_sparse(W: torch.Tensor):
    """torch_geometric.utils.dense_to_sparse (call site simdata_utils.py:519):
    row-major non-zeros, edge_index[0] = row, edge_index[1] = column."""
    idx = W.nonzero().t().contiguous()
    return idx, W[idx[0], idx[1]]


class BatchShim:
    """Stand-in for torch_geometric.data.Batch: the reference only reads
    .ptr and .batch (comb_opt.py:248,262), optionally .edge_index / .x."""

    def __init__(self, ptr, batch, edge_index=None, x=None):
        self.ptr = ptr
        self.batch = batch
        self.edge_index = edge_index
        self.x = x
        self.num_graphs = int(ptr.numel() - 1)

    def to(self, device):
        f = lambda t: None if t is None else t.to(device)
        return BatchShim(f(self.ptr), f(self.batch), f(self.edge_index), f(self.x))


def batch_from_sizes(sizes, edge_indices=None, xs=None) -> BatchShim:
    """Batch.from_data_list semantics (comb_opt.py:299,344): x concatenated on
    dim 0, edge_index offset by cumulative node counts, batch[i] = graph id,
    ptr = exclusive cumsum of node counts."""
    sizes = [int(s) for s in sizes]
    ptr = torch.zeros(len(sizes) + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(torch.tensor(sizes, dtype=torch.int64), 0)
    batch = torch.repeat_interleave(torch.arange(len(sizes), dtype=torch.int64), torch.tensor(sizes, dtype=torch.int64))
    ei = None
    if edge_indices is not None:
        ei = torch.cat([e.to(torch.int64) + ptr[g] for g, e in enumerate(edge_indices)], dim=1)
    x = None if xs is None else torch.cat(list(xs), dim=0)
    return BatchShim(ptr, batch, ei, x)


# ----------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------
def make_qnet_params(node_dim: int, emb_dim: int, seed: int = 0, scale: float = 1.0, dtype=torch.float32) -> dict:
    """Parameters with default nn.Linear init created in the order of
    QNet.__init__ (comb_opt.py:204-212) so that torch.manual_seed(seed) gives the
    same values as the reference constructor."""
    torch.manual_seed(seed)
    p = emb_dim
    shapes = [("theta1a", node_dim, p), ("theta1b", p, p), ("theta2", p, p), ("theta3", p, p),
              ("theta4", 1, p), ("theta5", 2 * p, 1), ("theta6", p, p), ("theta7", p, p)]
    out = {}
    for name, i, o in shapes:
        lin = nn.Linear(i, o, True, dtype=torch.float32)
        out[name + ".weight"] = (lin.weight.detach() * scale).to(dtype)
        out[name + ".bias"] = (lin.bias.detach() * scale).to(dtype)
    return out


def _lin(P, name, x):
    return F.linear(x, P[name + ".weight"], P[name + ".bias"])


# ----------------------------------------------------------------------------
# DQN Q-network, dense (reference arithmetic)
# ----------------------------------------------------------------------------
def qnet_forward_dense(P: dict, xv: torch.Tensor, Ws: torch.Tensor, ptr: torch.Tensor, batch: torch.Tensor,
                       T: int, norm_agg: bool = True, return_mu: bool = False):
    """comb_opt.py:217-275.  xv (B,Npad,F), Ws (B,Npad,Npad) -> q (sum N,)."""
    num_nodes_padded = xv.shape[1]
    batch_size = xv.shape[0]
    emb_dim = P["theta2.weight"].shape[0]
    conn_matrices = Ws
    mu = torch.zeros(batch_size, num_nodes_padded, emb_dim, dtype=xv.dtype)
    s1 = _lin(P, "theta1b", F.relu(_lin(P, "theta1a", xv)))                      # :232
    s3_1 = F.relu(_lin(P, "theta4", Ws.unsqueeze(3)))                            # :236
    s3_2 = torch.sum(s3_1, dim=1)                                                # :237
    s3 = _lin(P, "theta3", s3_2)                                                 # :238
    for _ in range(T):                                                           # :240-242
        s2 = _lin(P, "theta2", conn_matrices.matmul(mu))
        mu = F.relu(s1 + s2 + s3)
    sizes = (ptr[1:] - ptr[:-1]).tolist()                                        # :248-249
    select = []
    for n in sizes:                                                              # :254-256
        select += [True] * n + [False] * (num_nodes_padded - n)
    mu_raw = mu.reshape(-1, emb_dim)[torch.tensor(select, dtype=torch.bool)]     # :259
    if norm_agg:                                                                 # :261-268
        pool = scatter_mean(mu_raw, batch, dim_size=batch_size)
    else:
        pool = scatter_add(mu_raw, batch, dim_size=batch_size)
    pool_expanded = torch.repeat_interleave(pool, torch.tensor(sizes, dtype=torch.int64), dim=0)
    global_state = _lin(P, "theta6", pool_expanded)
    local_action = _lin(P, "theta7", mu_raw)                                     # :270
    out = F.relu(torch.cat([global_state, local_action], dim=1))                 # :272
    out = _lin(P, "theta5", out).squeeze()                                       # :273
    if return_mu:
        return out, mu_raw
    return out


def greedy_action(estimated_rewards: np.ndarray, reachable_nodes):
    """comb_opt.py:319-326: index into the (ascending) neighbour list, node id, value."""
    reachable_rewards = np.asarray(estimated_rewards)[reachable_nodes]
    action = int(np.argmax(reachable_rewards))
    return action, int(reachable_nodes[action]), reachable_rewards[action]


def dqn_loss(q: torch.Tensor, actions, ptr: torch.Tensor, targets) -> torch.Tensor:
    """comb_opt.py:355-359: SmoothL1(mean) of q[a + ptr[:-1]] vs targets."""
    sel = torch.as_tensor(actions, dtype=torch.int64) + ptr[:-1]
    return F.smooth_l1_loss(q[sel], torch.as_tensor(targets, dtype=q.dtype))


# ----------------------------------------------------------------------------
# sparse restatement (what the CUDA kernels compute)
# ----------------------------------------------------------------------------
def in_degree(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """c_j = #{i : W[i,j] != 0} (column count; the theta4 sum runs over dim=1, comb_opt.py:237)."""
    return torch.zeros(num_nodes, dtype=torch.int64).index_add_(
        0, edge_index[1], torch.ones(edge_index.shape[1], dtype=torch.int64))


def s2v_embed_sparse(P: dict, x: torch.Tensor, edge_index: torch.Tensor, n_pad_node: torch.Tensor, T: int,
                     return_all: bool = False):
    """Embedding rounds on the unpadded node set.
    x (sumN,F); edge_index (2,sumE) global ids, [0]=row i (gathers), [1]=column j;
    n_pad_node (sumN,) padded size of the graph each node belongs to."""
    dt = x.dtype
    n = x.shape[0]
    p = P["theta2.weight"].shape[0]
    s1 = _lin(P, "theta1b", F.relu(_lin(P, "theta1a", x)))
    w4 = P["theta4.weight"].reshape(-1)
    b4 = P["theta4.bias"]
    r1 = F.relu(w4 + b4)
    r0 = F.relu(b4)
    c = in_degree(edge_index, n).to(dt)
    z = c[:, None] * r1[None, :] + (n_pad_node.to(dt) - c)[:, None] * r0[None, :]
    s3 = _lin(P, "theta3", z)
    mu = torch.zeros(n, p, dtype=dt)
    mus = []
    for _ in range(T):
        agg = torch.zeros(n, p, dtype=dt).index_add_(0, edge_index[0], mu[edge_index[1]])
        mu = F.relu(s1 + _lin(P, "theta2", agg) + s3)
        mus.append(mu)
    if return_all:
        return mu, mus
    return mu


def s2v_embed_sparse_masked(P: dict, x: torch.Tensor, edge_index: torch.Tensor, n_pad_node: torch.Tensor, T: int, masks):
    """Same arithmetic as s2v_embed_sparse with the ReLU DECISIONS of the T rounds prescribed: mu^t = z^t * masks[t]
    (masks[t] bool [sumN, p]).  The gradient of a ReLU network is discontinuous where a pre-activation crosses zero; an
    implementation that rounds differently may decide an element within fp32 resolution of the kink the other way, and
    a loss that touches every node then routes that element's whole contribution differently.  Holding the decisions
    fixed makes the comparison well posed: gradients must agree to rounding, and the decisions themselves are compared
    separately (count of elements with (z^t > 0) != masks[t], each of which must sit at the kink).
    Returns (mu^T, [z^1..z^T])."""
    dt = x.dtype
    n = x.shape[0]
    p = P["theta2.weight"].shape[0]
    s1 = _lin(P, "theta1b", F.relu(_lin(P, "theta1a", x)))
    w4 = P["theta4.weight"].reshape(-1)
    b4 = P["theta4.bias"]
    r1 = F.relu(w4 + b4)
    r0 = F.relu(b4)
    c = in_degree(edge_index, n).to(dt)
    z3 = c[:, None] * r1[None, :] + (n_pad_node.to(dt) - c)[:, None] * r0[None, :]
    s3 = _lin(P, "theta3", z3)
    mu = torch.zeros(n, p, dtype=dt)
    zs = []
    for t in range(T):
        agg = torch.zeros(n, p, dtype=dt).index_add_(0, edge_index[0], mu[edge_index[1]])
        z = s1 + _lin(P, "theta2", agg) + s3
        zs.append(z)
        mu = z * masks[t].to(dt)
    return mu, zs


def head_sparse(P: dict, mu: torch.Tensor, ptr: torch.Tensor, batch: torch.Tensor, norm_agg: bool = True,
                names=("theta5", "theta6", "theta7")) -> torch.Tensor:
    t5, t6, t7 = names
    B = ptr.numel() - 1
    pool = scatter_mean(mu, batch, B) if norm_agg else scatter_add(mu, batch, B)
    G = _lin(P, t6, pool)
    L = _lin(P, t7, mu)
    rep = F.relu(torch.cat([G[batch], L], dim=1))
    return _lin(P, t5, rep).squeeze(-1)


def qnet_forward_sparse(P: dict, x: torch.Tensor, edge_index: torch.Tensor, ptr: torch.Tensor, n_pad,
                        T: int, norm_agg: bool = True) -> torch.Tensor:
    """Sparse form of comb_opt.py:217-275.  n_pad: int or (B,) tensor."""
    B = ptr.numel() - 1
    sizes = ptr[1:] - ptr[:-1]
    batch = torch.repeat_interleave(torch.arange(B, dtype=torch.int64), sizes)
    n_pad_t = torch.as_tensor(n_pad, dtype=torch.int64)
    if n_pad_t.dim() == 0:
        n_pad_t = n_pad_t.expand(B)
    mu = s2v_embed_sparse(P, x, edge_index, n_pad_t[batch], T)
    return head_sparse(P, mu, ptr, batch, norm_agg)


# ----------------------------------------------------------------------------
# PPO-GNN-LSTM s2v path, dense (reference arithmetic)
# ----------------------------------------------------------------------------
def ppo_s2v_dense(P: dict, xv: torch.Tensor, Ws: torch.Tensor, T: int) -> torch.Tensor:
    """models_basic_lstm.py:479-505.  xv (S,N,F), Ws (N,N) shared -> mu (S,N,p)."""
    num_nodes = xv.shape[1]
    batch_size = xv.shape[0]
    p = P["theta2.weight"].shape[0]
    mu = torch.zeros(batch_size, num_nodes, p, dtype=xv.dtype)
    s1 = _lin(P, "theta1b", F.relu(_lin(P, "theta1a", xv)))                      # :493
    s3_1 = F.relu(_lin(P, "theta4", Ws.unsqueeze(2)))                            # :497
    s3_2 = torch.sum(s3_1, dim=0)                                                # :498
    s3 = _lin(P, "theta3", s3_2)                                                 # :499
    for _ in range(T):                                                           # :501-503
        s2 = _lin(P, "theta2", Ws.matmul(mu))
        mu = F.relu(s1 + s2 + s3)
    return mu


def ppo_embeddings_dense(P, nfm, ei, T, lstm=None, hidden=None):
    """models_basic_lstm.py:507-524 with qnet == 's2v'."""
    Ws = to_dense_adj(ei)[0].to(nfm.dtype)
    mu = ppo_s2v_dense(P, nfm, Ws, T)
    mu = mu.reshape(nfm.shape[0], -1, mu.shape[-1])
    if lstm is not None:
        mu, hidden = lstm(mu, hidden)
    return mu, hidden


def _ppo_head(P, mu, suffix):
    seq_len, num_nodes, p = mu.shape
    mu_meanpool = mu.mean(dim=1, keepdim=True)                                   # :530
    expander = torch.tensor([num_nodes] * seq_len, dtype=torch.int64)
    expanded = torch.repeat_interleave(mu_meanpool, expander, dim=0).reshape(seq_len, num_nodes, p)
    global_state = _lin(P, "theta6" + suffix, expanded)
    local_action = _lin(P, "theta7" + suffix, mu)
    rep = F.relu(torch.cat((global_state, local_action), dim=2))
    return _lin(P, "theta5" + suffix, rep).squeeze(-1)


def ppo_pi_dense(P, nfm, ei, reachable, T, lstm=None, hidden=None):
    """models_basic_lstm.py:526-542: masked softmax over nodes."""
    if nfm.dim() == 2:
        nfm, reachable = nfm[None], reachable[None]
    mu, hidden = ppo_embeddings_dense(P, nfm, ei, T, lstm, hidden)
    logits = _ppo_head(P, mu, "_pi").clone()
    logits[~reachable] = -torch.inf                                              # :539
    return F.softmax(logits, dim=1), hidden


def ppo_v_dense(P, nfm, ei, reachable, T, critic="q", lstm=None, hidden=None):
    """models_basic_lstm.py:544-566: critic 'q' = max over reachable, 'v' = linear on pool."""
    if nfm.dim() == 2:
        nfm, reachable = nfm[None], reachable[None]
    mu, hidden = ppo_embeddings_dense(P, nfm, ei, T, lstm, hidden)
    if critic == "v":
        v = _lin(P, "theta5_v", mu.mean(dim=1, keepdim=True)).squeeze(-1)        # :552
    else:
        qvals = _ppo_head(P, mu, "_v").clone()
        qvals[~reachable] = -torch.inf                                           # :563
        v = qvals.max(dim=1)[0].unsqueeze(-1)                                    # :564
    return v, hidden


# ----------------------------------------------------------------------------
# integer work: CSR layout, masks (bit-exact targets)
# ----------------------------------------------------------------------------
def csr_from_edge_index(edge_index: np.ndarray, num_nodes: int):
    """Deterministic CSR of the row form (row i lists its columns j in edge
    order) and of the transpose (column j lists its rows i in edge order).
    Stable sorts, so edge order inside a row is the input order; on the
    reference's row-major EI (simdata_utils.py:519) the row form is the identity
    permutation.  Returns int32 arrays:
    rowptr (n+1), col (E), rowptr_t (n+1), col_t (E), indeg (n)."""
    ei = np.asarray(edge_index)
    src, dst = ei[0].astype(np.int64), ei[1].astype(np.int64)
    order = np.argsort(src, kind="stable")
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.add.at(rowptr, src + 1, 1)
    rowptr = np.cumsum(rowptr)
    col = dst[order]
    order_t = np.argsort(dst, kind="stable")
    rowptr_t = np.zeros(num_nodes + 1, dtype=np.int64)
    np.add.at(rowptr_t, dst + 1, 1)
    indeg = rowptr_t[1:].copy()
    rowptr_t = np.cumsum(rowptr_t)
    col_t = src[order_t]
    return (rowptr.astype(np.int32), col.astype(np.int32), rowptr_t.astype(np.int32),
            col_t.astype(np.int32), indeg.astype(np.int32))


def masked_argmax_per_graph(q: np.ndarray, mask: np.ndarray, ptr: np.ndarray):
    """Per graph: (local node id of the first maximum among mask==True, value).
    Equals np.argmax over the ascending neighbour list (comb_opt.py:319-326) and
    torch argmax/max after logits[~reachable] = -inf (models_basic_lstm.py:539,563;
    rl_policy.py:536).  Graphs with an empty mask give (-1, -inf)."""
    B = len(ptr) - 1
    arg = np.full(B, -1, dtype=np.int64)
    val = np.full(B, -np.inf, dtype=q.dtype)
    for g in range(B):
        lo, hi = int(ptr[g]), int(ptr[g + 1])
        idx = np.nonzero(mask[lo:hi])[0]
        if idx.size:
            k = int(np.argmax(q[lo:hi][idx]))
            arg[g] = idx[k]
            val[g] = q[lo + idx[k]]
    return arg, val


def masked_softmax_per_graph(logits: torch.Tensor, mask: torch.Tensor, ptr: torch.Tensor) -> torch.Tensor:
    """Per graph softmax after logits[~mask] = -inf (models_basic_lstm.py:539-540)."""
    out = torch.zeros_like(logits)
    for g in range(ptr.numel() - 1):
        lo, hi = int(ptr[g]), int(ptr[g + 1])
        l = logits[lo:hi].clone()
        l[~mask[lo:hi]] = -torch.inf
        out[lo:hi] = F.softmax(l, dim=0)
    return out
