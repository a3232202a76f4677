"""Drop-in for the batched variable-size PPO policy of the reference's custom PPO stack.

Mirrors modules/ppo/models_ngo.py: ``FeatureExtractor`` (233-300: flat observation vectors -> PyG batch -> GATv2
with share_weights=True, heads 2, 5 layers -> node features [mu | batch | reachable | num_nodes]), ``Actor``
(387-486: per-node LSTM state, scatter_mean pool, theta6/7/5 head, -inf mask, logits padded per graph to
``hp.max_possible_nodes`` for ``Categorical``), ``Critic`` (501-588: same head, ``scatter_max`` over reachable nodes, or
theta8 on the pool) and ``MaskablePPOPolicy`` (125-160).  Same constructor arguments, parameter names and return
shapes; ``hidden_cell`` bookkeeping (per-node LSTM state slots, ``selector``) is the reference's and stays torch /
cuDNN like every LSTM of this path.

On the kernels: the GATv2 layers (csrc/s2v_gat.cu), per-graph pooling + head (ops.s2v_head, mean pool over the
variable-size graphs of every time step), masked per-graph max (ops.masked_max).  The flat observations are
de-serialised with ONE device-to-host copy of the five size fields per observation instead of a ``.tolist()`` per
observation (models_ngo.py:256), everything else is index arithmetic on the device.  CUDA only.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch import distributions

from . import ops
from ._lib import S2VError
from .gat import GATv2, gat_layout_from_batch
from .layout import GraphLayout


class FeatureExtractor(nn.Module):
    """models_ngo.py:233-300."""

    def __init__(self, state_dim, hp):
        super().__init__()
        self.hp = hp
        self.emb_dim = hp.emb_dim
        self.T = 5
        self.node_dim = hp.node_dim
        self.gat = GATv2(in_channels=self.node_dim, hidden_channels=self.emb_dim, heads=2, num_layers=self.T,
                         out_channels=self.emb_dim, share_weights=True, concat=True)

    def forward(self, state):
        """state f32[seq_len, bsize, flatvecdim], each row
        [nfm (max_nodes*F) | edge_list (2*max_edges) | reachable (max_nodes) | num_nodes, max_nodes, num_edges, max_edges, node_dim]
        -> (features f32[seq_len, nodes_in_batch / seq_len, emb_dim + 3], nodes_in_batch, None, num_nodes)."""
        assert state.dim() == 3
        if not state.is_cuda:
            raise S2VError("simmodel_b200.ngo.FeatureExtractor runs on CUDA only (no CPU fallback)")
        seq_len, batch_size, flat = state.shape
        dev = state.device
        obs = state.reshape(-1, flat)
        G = obs.shape[0]
        sizes = obs[:, -5:].to(torch.int64).cpu()                      # one D2H copy for all observations
        n_i, max_nodes, e_i, max_edges, node_dim = (sizes[:, k] for k in range(5))
        max_nodes, max_edges, node_dim = int(max_nodes[0]), int(max_edges[0]), int(node_dim[0])
        assert bool((sizes[:, 1] == max_nodes).all()) and bool((sizes[:, 3] == max_edges).all())
        n_dev, e_dev = n_i.to(dev), e_i.to(dev)
        ptr = torch.zeros(G + 1, dtype=torch.int64, device=dev)
        ptr[1:] = torch.cumsum(n_dev, 0)
        total = int(ptr[-1])
        gid = torch.repeat_interleave(torch.arange(G, device=dev), n_dev)
        local = torch.arange(total, device=dev) - ptr[gid]
        nfm = obs[:, :node_dim * max_nodes].reshape(G, max_nodes, node_dim)
        x = nfm[gid, local].contiguous()                                # first num_nodes rows of every observation
        reach = obs[:, node_dim * max_nodes + 2 * max_edges: node_dim * max_nodes + 2 * max_edges + max_nodes]
        reachable = reach[gid, local].to(torch.int64).to(torch.float32)
        eptr = torch.zeros(G + 1, dtype=torch.int64, device=dev)
        eptr[1:] = torch.cumsum(e_dev, 0)
        egid = torch.repeat_interleave(torch.arange(G, device=dev), e_dev)
        elocal = torch.arange(int(eptr[-1]), device=dev) - eptr[egid]
        el = obs[:, node_dim * max_nodes: node_dim * max_nodes + 2 * max_edges].reshape(G, 2, max_edges)
        ei = el[egid, :, elocal].t().to(torch.int64) + ptr[egid]        # first num_edges columns, offset per graph
        layout = gat_layout_from_batch(ei.contiguous(), ptr, gid)
        emb = self.gat.forward_layout(x, layout)                        # [total, emb_dim]
        # Feature rows = [embedding | graph id within one step | reachable flag | node-count column].  The reference
        # fills the node-count column with the first rows-per-step entries of (per-graph node counts, zero padded to
        # one entry per node) and tiles columns 2 and 4 over the steps (models_ngo.py:286-298); the consumers read
        # only its first batch_size entries of step 0, the same values here.
        per_step = total // seq_len
        counts = torch.zeros(total, dtype=torch.float32, device=dev)
        counts[:G] = n_dev
        feat = torch.empty(seq_len, per_step, self.emb_dim + 3, dtype=torch.float32, device=dev)
        feat[:, :, :self.emb_dim] = emb.view(seq_len, per_step, self.emb_dim)
        feat[:, :, self.emb_dim] = gid[:per_step].to(torch.float32)
        feat[:, :, self.emb_dim + 1] = reachable.view(seq_len, per_step)
        feat[:, :, self.emb_dim + 2] = counts[:per_step]
        return feat, total, None, counts[:per_step].repeat(seq_len)


class _HeadBase(nn.Module):
    """Shared bookkeeping of Actor / Critic: feature split, per-node LSTM state, the layout of seq_len x B graphs."""

    def reset_init_state(self, batch_size, device):
        self.hidden_cell = (torch.zeros(self.hp.recurrent_layers, batch_size, self.hidden_size).to(device),
                            torch.zeros(self.hp.recurrent_layers, batch_size, self.hidden_size).to(device))

    def _prepare(self, features, terminal, selector):
        assert features.dim() == 3
        if not features.is_cuda:
            raise S2VError("simmodel_b200.ngo heads run on CUDA only (no CPU fallback)")
        seq_len, num_nodes, _ = features.shape
        max_nodes = self.hp.max_possible_nodes
        dev = features.device
        mu_raw, batch, reachable, num_nodes_list = torch.split(features, [self.emb_dim, 1, 1, 1], dim=2)
        batch_size = int(batch.max()) + 1
        batch = batch.squeeze(-1).to(torch.int64)
        reachable = reachable.squeeze(-1).bool()
        if self.hidden_cell is None or max_nodes * batch_size != self.hidden_cell[0].shape[1]:
            self.reset_init_state(max_nodes * batch_size, dev)
        if terminal is not None:                  # state slots of finished episodes must have been cleared by the caller
            done_slots = terminal.to(torch.bool).reshape(batch_size, 1).expand(batch_size, max_nodes).reshape(-1)
            for state in self.hidden_cell:
                assert float(state[:, done_slots, :].abs().sum()) == 0
        sizes = num_nodes_list[0].squeeze(-1)[:batch_size].to(torch.int64)             # nodes per graph
        if self.lstm_on:
            if selector is None:                  # slot b*max_nodes + i belongs to node i of graph b, i < its node count
                selector = (torch.arange(max_nodes)[None, :] < sizes.cpu()[:, None]).reshape(-1)
            full_output, out_hc = self.lstm(mu_raw.contiguous(), (self.hidden_cell[0][:, selector, :], self.hidden_cell[1][:, selector, :]))
            self.hidden_cell[0][:, selector, :] = out_hc[0]
            self.hidden_cell[1][:, selector, :] = out_hc[1]
            mu_mem = full_output
        else:
            mu_mem = mu_raw
        if mu_mem.shape[-1] != self.emb_dim:
            raise NotImplementedError("the head kernels need hidden_size == emb_dim (got %d vs %d)" % (mu_mem.shape[-1], self.emb_dim))
        # seq_len x B graphs: graph (s, b) owns rows s*num_nodes + [ptr_b, ptr_b+1)
        ptr1 = torch.zeros(batch_size + 1, dtype=torch.int64, device=dev)
        ptr1[1:] = torch.cumsum(sizes, 0)
        ptr = (torch.arange(seq_len, device=dev)[:, None] * num_nodes + ptr1[None, :-1]).reshape(-1)
        ptr = torch.cat((ptr, torch.tensor([seq_len * num_nodes], device=dev)))
        gid = (torch.arange(seq_len, device=dev)[:, None] * batch_size + batch[0][None, :]).reshape(-1)
        layout = GraphLayout.from_batch(torch.zeros(2, 0, dtype=torch.int64, device=dev), ptr, max_nodes, gid)
        flat = mu_mem.reshape(seq_len * num_nodes, self.emb_dim).contiguous()
        return flat, layout, reachable.reshape(-1), batch, sizes, seq_len, num_nodes, batch_size, max_nodes


class Actor(_HeadBase):
    """models_ngo.py:387-486."""

    def __init__(self, state_dim, action_dim, continuous_action_space, trainable_std_dev, init_log_std_dev=None, hp=None):
        super().__init__()
        self.lstm_on = hp.lstm_on
        self.hp = hp
        self.emb_dim = hp.emb_dim
        self.action_dim = action_dim
        self.continuous_action_space = continuous_action_space
        self.hidden_size = hp.hidden_size
        self.num_recurrent_layers = hp.recurrent_layers
        self.theta5_pi = nn.Linear(2 * self.emb_dim, 1, True)
        if self.lstm_on:
            self.lstm = nn.LSTM(self.emb_dim, self.hidden_size, num_layers=self.num_recurrent_layers)
            self.theta6_pi = nn.Linear(self.hidden_size, self.emb_dim, True)
            self.theta7_pi = nn.Linear(self.hidden_size, self.emb_dim, True)
        else:
            self.theta6_pi = nn.Linear(self.emb_dim, self.emb_dim, True)
            self.theta7_pi = nn.Linear(self.emb_dim, self.emb_dim, True)
        self.log_std_dev = nn.Parameter(init_log_std_dev * torch.ones(33, dtype=torch.float), requires_grad=trainable_std_dev)
        self.covariance_eye = torch.eye(33).unsqueeze(0)
        self.hidden_cell = None

    def forward(self, features, terminal=None, selector=None):
        """-> Categorical over logits f32[seq_len, B, max_possible_nodes] (unreachable / padded entries -inf), on the CPU
        like the reference's (models_ngo.py:484)."""
        flat, layout, reach, batch, sizes, seq_len, num_nodes, B, max_nodes = self._prepare(features, terminal, selector)
        w = [t for n in ("theta5_pi", "theta6_pi", "theta7_pi") for t in (getattr(self, n).weight, getattr(self, n).bias)]
        logits = ops.s2v_head(layout, flat, w, True)                                   # scatter_mean pool + head
        logits = logits.masked_fill(~reach, float("-inf")).view(seq_len, num_nodes)
        # pad every graph to max_nodes (models_ngo.py:468-479)
        ptr1 = torch.zeros(B + 1, dtype=torch.int64, device=flat.device)
        ptr1[1:] = torch.cumsum(sizes, 0)
        local = torch.arange(num_nodes, device=flat.device) - ptr1[batch[0]]
        padded = torch.full((seq_len, B, max_nodes), float("-inf"), dtype=logits.dtype, device=flat.device)
        padded[:, batch[0], local] = logits
        if self.continuous_action_space:
            assert False  # not implemented in the reference either
        return distributions.Categorical(logits=padded.to("cpu"))

    def numTrainableParameters(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)


class Critic(_HeadBase):
    """models_ngo.py:501-588."""

    def __init__(self, state_dim, hp, lstm=None):
        super().__init__()
        self.lstm_on = hp.lstm_on
        self.emb_dim = hp.emb_dim
        self.hidden_size = hp.hidden_size
        self.num_recurrent_layers = hp.recurrent_layers
        self.hp = hp
        assert hp.critic in ["q", "v"]
        if self.lstm_on:
            self.lstm = nn.LSTM(self.emb_dim, self.hidden_size, num_layers=self.num_recurrent_layers) if lstm is None else lstm
            if hp.critic == "q":
                self.theta6_v = nn.Linear(self.hidden_size, self.emb_dim, True)
                self.theta7_v = nn.Linear(self.hidden_size, self.emb_dim, True)
                self.theta5_v = nn.Linear(2 * self.emb_dim, 1, True)
        elif hp.critic == "q":
            self.theta6_v = nn.Linear(self.emb_dim, self.emb_dim, True)
            self.theta7_v = nn.Linear(self.emb_dim, self.emb_dim, True)
            self.theta5_v = nn.Linear(2 * self.emb_dim, 1, True)
        if hp.critic == "v":
            self.theta8_v = nn.Linear(self.emb_dim, 1, True)
        self.hidden_cell = None

    def forward(self, features, terminal=None, selector=None):
        """-> v f32[seq_len, B, 1]."""
        flat, layout, reach, batch, sizes, seq_len, num_nodes, B, max_nodes = self._prepare(features, terminal, selector)
        if self.hp.critic == "q":
            w = [t for n in ("theta5_v", "theta6_v", "theta7_v") for t in (getattr(self, n).weight, getattr(self, n).bias)]
            qvals = ops.s2v_head(layout, flat, w, True)
            v, _ = ops.masked_max(layout, qvals, reach)                                # qvals[~reachable] = -inf; scatter_max
            v = v.view(seq_len, B)
        else:
            pool = ops.segment_pool(layout, flat, True).view(seq_len, B, self.emb_dim)
            v = torch.nn.functional.linear(pool, self.theta8_v.weight, self.theta8_v.bias).squeeze(-1)
        return v.unsqueeze(-1)


class MaskablePPOPolicy(nn.Module):
    """models_ngo.py:125-160."""

    def __init__(self, state_dim, action_dim, continuous_action_space, trainable_std_dev, init_log_std_dev=None, hp=None, verbose=False):
        super().__init__()
        self.description = ("Action Masked PPO Policy with A+C LSTMs and GATv2 feature extraction (B200 kernels)" if hp.lstm_on
                            else "Action Masked PPO Policy with LSTM switched off and GATv2 feature extraction (B200 kernels)")
        self.action_dim = action_dim
        self.continuous_action_space = continuous_action_space
        self.hp = hp
        self.FE = FeatureExtractor(state_dim, hp)
        self.PI = Actor(state_dim, action_dim, continuous_action_space, trainable_std_dev, init_log_std_dev, hp=hp)
        self.V = Critic(state_dim, hp)
        if verbose:
            print(self.numTrainableParameters()[1])

    def forward(self, features, terminal=None, selector=None):
        return self.PI(features, terminal, selector), self.V(features, terminal, selector)

    def get_values(self, features, terminal=None, selector=None):
        return self.V(features, terminal, selector)

    def numTrainableParameters(self):
        ps = self.description + "\n------------------------------------------\n"
        total = 0
        for name, p in self.named_parameters():
            if p.requires_grad:
                total += int(np.prod(p.shape))
            ps += "{:24s} {:12s} requires_grad={}\n".format(name, str(list(p.shape)), p.requires_grad)
        ps += "Total number of trainable parameters: {}\n".format(total)
        ps += "------------------------------------------"
        assert total == sum(p.numel() for p in self.parameters() if p.requires_grad)
        return total, ps
