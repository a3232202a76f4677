"""Generate tests/golden/ngo_*.npz by running the UNMODIFIED reference classes of modules/ppo/models_ngo.py
(TEST INFRASTRUCTURE ONLY; build container only:  python -m oracle.gen_golden_ngo).

MaskablePPOPolicy = FeatureExtractor (233-300) + Actor (387-486) + Critic (501-588) import from /root/reference as they
are.  Absent third-party pieces are restated: torch_scatter scatter_mean / scatter_max along dim 1 (below), PyG Data /
Batch / GATv2Conv / BasicGNN (oracle/gat_ref.py; the conv arithmetic is unpinned against real PyG, see there).
"""
from __future__ import annotations

import contextlib
import importlib
import io
import types

import numpy as np
import torch

from oracle import gat_ref, ref_import
from oracle.gen_golden import _np, _save, _scale_of
from simmodel_b200 import graphs


def scatter_mean_dim1(src, index, dim=1):
    """torch_scatter.scatter_mean(src [S,N,p], index [S,N], dim=1) -> [S,B,p] (count clamped >= 1)."""
    assert dim == 1
    S, N, p = src.shape
    B = int(index.max()) + 1
    out = torch.zeros(S, B, p, dtype=src.dtype).scatter_add(1, index.unsqueeze(-1).expand(-1, -1, p), src)
    cnt = torch.zeros(S, B, dtype=src.dtype).scatter_add(1, index, torch.ones(S, N, dtype=src.dtype)).clamp(min=1)
    return out / cnt.unsqueeze(-1)


def scatter_max_dim1(src, index, dim=1):
    """torch_scatter.scatter_max(src [S,N], index [S,N], dim=1) -> (values [S,B], argmax [S,B])."""
    assert dim == 1
    S, N = src.shape
    B = int(index.max()) + 1
    out = torch.full((S, B), float("-inf"), dtype=src.dtype).scatter_reduce(1, index, src, "amax", include_self=True)
    arg = torch.zeros(S, B, dtype=torch.int64)
    for s in range(S):
        for b in range(B):
            cand = ((index[s] == b) & (src[s] == out[s, b])).nonzero()
            arg[s, b] = int(cand[0]) if cand.numel() else N
    return out, arg


def load_ngo_reference():
    ref_import.load_reference()
    m = importlib.import_module("modules.ppo.models_ngo")
    m.scatter_mean, m.scatter_max = scatter_mean_dim1, scatter_max_dim1
    m.Data = gat_ref.Data

    class _Batch:
        from_data_list = staticmethod(gat_ref.batch_from_data_list)
    m.Batch = _Batch
    return m


def flat_obs(x, ei, reach, max_nodes, max_edges):
    n, F = x.shape
    E = ei.shape[1]
    nfm = torch.zeros(max_nodes, F)
    nfm[:n] = x
    el = torch.zeros(2, max_edges)
    el[:, :E] = ei.to(torch.float32)
    r = torch.zeros(max_nodes)
    r[reach] = 1.0
    return torch.cat((nfm.reshape(-1), el.reshape(-1), r, torch.tensor([n, max_nodes, E, max_edges, F], dtype=torch.float32)))


def case(m, name, topos, seq_len, lstm_on, critic, seed):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    max_nodes = max(t.num_nodes for t in topos)
    max_edges = max(t.num_edges for t in topos)
    hp = types.SimpleNamespace(emb_dim=64, node_dim=graphs.NODE_DIM, lstm_on=lstm_on, hidden_size=64, recurrent_layers=1,
                               max_possible_nodes=max_nodes, critic=critic)
    with contextlib.redirect_stdout(io.StringIO()):
        pol = m.MaskablePPOPolicy(None, max_nodes, False, False, init_log_std_dev=0.0, hp=hp)
    with torch.no_grad():
        for k, prm in pol.named_parameters():
            if k.endswith(".bias") and "convs" in k and "lin_" not in k:
                prm.copy_(torch.randn(prm.shape) * 0.1)
    P = {k: v.detach().clone() for k, v in pol.state_dict().items()}
    B = len(topos)
    obs = []
    for s in range(seq_len):
        for t in topos:
            x = graphs.synth_nfm(t, 1, gen)[0]
            cur = int(x[:, 1].argmax())
            obs.append(flat_obs(x, t.edge_index, t.neighbors(cur) or [cur], max_nodes, max_edges))
    state = torch.stack(obs).reshape(seq_len, B, -1)
    feats, nodes_in_batch, _, num_nodes = pol.FE(state)
    dist, v = pol(feats)
    wp = torch.randn(seq_len, B, max_nodes, generator=gen)
    wv = torch.randn(seq_len, B, 1, generator=gen)
    pol.zero_grad()
    ((dist.probs * wp).sum() + (v * wv).sum()).backward()
    grads = {k: (prm.grad.detach().clone() if prm.grad is not None else torch.zeros_like(prm)) for k, prm in pol.named_parameters()}
    h_pi = None if not lstm_on else [t.detach().clone() for t in pol.PI.hidden_cell]
    print("  %-26s features %s  probs max %.3g  v scale %.3g" % (name, tuple(feats.shape), float(dist.probs.max()), _scale_of(v)))
    meta = dict(kind="ngo", seq_len=seq_len, B=B, max_nodes=max_nodes, max_edges=max_edges, lstm_on=lstm_on, critic=critic, seed=seed,
                nodes_in_batch=int(nodes_in_batch))
    out = dict(features=_np(feats), probs=_np(dist.probs), v=_np(v), num_nodes=_np(num_nodes))
    if h_pi is not None:
        out["h_pi"], out["c_pi"] = _np(h_pi[0]), _np(h_pi[1])
    _save(name, meta, P={k: _np(v_) for k, v_ in P.items()}, inp=dict(state=_np(state), wp=_np(wp), wv=_np(wv)), out=out,
          grad={k: _np(g) for k, g in grads.items()})


def main():
    m = load_ngo_reference()
    print("models_ngo cases (MaskablePPOPolicy = FeatureExtractor + Actor + Critic)")
    m3, m5, mem = graphs.load_topology("Manhattan3"), graphs.load_topology("Manhattan5"), graphs.load_topology("MemTask")
    case(m, "ngo_m3m5_seq3_nolstm_q", [m3, m5], 3, False, "q", seed=50)
    case(m, "ngo_m5m3mem_seq2_lstm_q", [m5, m3, mem], 2, True, "q", seed=51)
    case(m, "ngo_m3m5_seq1_nolstm_v", [m3, m5], 1, False, "v", seed=52)


if __name__ == "__main__":
    main()
