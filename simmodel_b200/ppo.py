"""Drop-in for the s2v branch of the reference's PPO-GNN-LSTM policy.

Mirrors modules/ppo/models_basic_lstm.py:340-566 (PPO_GNN_Model /
PPO_GNN_Single_LSTM with config['qnet'] == 's2v'): same constructor arguments
(config, hp, tp), same parameter names (theta1a..theta4, theta{5,6,7}_{pi,v},
lstm.*), same ``s2v / create_embeddings / pi / v`` signatures and return shapes.
The GNN embedding, pooling, head, reachable-node mask, softmax and max run in the
sm_100a kernels; the nn.LSTM between embedding and head stays cuDNN (SURVEY.md
§2.2: out of scope), as do the PPO loss and the trainer loop (train_net / learn),
which call pi() and v() unchanged.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.optim as optim

from . import ops
from ._lib import S2VError
from .gat import GATv2, strip_self_loops
from .layout import GraphLayout


class _PPOModelBase(nn.Module):
    """Reference constructor / bookkeeping (models_basic_lstm.py:340-462); the kernel-backed
    methods come from S2VKernelMixin (PPO_GNN_Single_LSTM at the end of this file)."""

    def __init__(self, config, hp, tp, device="cuda"):
        super().__init__()
        self.hp, self.tp, self.config = hp, tp, config
        self.data, self.ei = [], []
        self.node_dim = hp.node_dim
        self.emb_dim = hp.emb_dim
        self.num_rollouts = hp.parallel_rollouts
        self.num_epochs = hp.batch_size
        assert config["lstm_type"] in ["None", "EMB", "FE"]
        assert config["critic"] in ["q", "v"]
        if config["qnet"] not in ("s2v", "gat2"):
            raise NotImplementedError("simmodel_b200 implements qnet='s2v' and qnet='gat2'")
        if tp is not None and tp.get("base_checkpoint_path"):
            Path(tp["base_checkpoint_path"]).mkdir(parents=True, exist_ok=True)
        self.description = "PPO policy, s2v extractor (B200 kernels), lstm, action masking"
        dev = torch.device(device)
        # creation order == models_basic_lstm.py:414-460
        if config["qnet"] == "gat2":
            self.description = "PPO policy, GATv2 extractor (B200 kernels), lstm, action masking"
            self.gat = GATv2(in_channels=self.node_dim, hidden_channels=self.emb_dim * 2, heads=config["gat_heads"],
                             num_layers=config["emb_iterT"], out_channels=self.emb_dim,
                             share_weights=config["gat_share_weights"], concat=config["gat_concat"]).to(dev)
            p = self.emb_dim
        else:
            self.emb_dim = config["emb_dim"]
            self.T = config["emb_iterT"]
            p = self.emb_dim
            self.theta1a = nn.Linear(self.node_dim, p, True, device=dev)
            self.theta1b = nn.Linear(p, p, True, device=dev)
            self.theta2 = nn.Linear(p, p, True, device=dev)
            self.theta3 = nn.Linear(p, p, True, device=dev)
            self.theta4 = nn.Linear(1, p, True, device=dev)
        if config["lstm_type"] == "EMB":
            self.lstm_on = True
            self.lstm = nn.LSTM(p, p, device=dev)
        elif config["lstm_type"] == "FE":
            self.lstm_on = True
            self.lstm = nn.LSTM(self.node_dim, self.node_dim, device=dev)
        else:
            self.lstm_on = False
        self.theta5_pi = nn.Linear(2 * p, 1, True, device=dev)
        self.theta6_pi = nn.Linear(p, p, True, device=dev)
        self.theta7_pi = nn.Linear(p, p, True, device=dev)
        if config["critic"] == "v":
            self.theta5_v = nn.Linear(p, 1, True, device=dev)
        else:
            self.theta5_v = nn.Linear(2 * p, 1, True, device=dev)
            self.theta6_v = nn.Linear(p, p, True, device=dev)
            self.theta7_v = nn.Linear(p, p, True, device=dev)
        self.optimizer = optim.Adam(self.parameters(), lr=self.hp.learning_rate)

    def put_data(self, transition_buffer, ei=None):
        self.data.append(transition_buffer)
        self.ei.append(ei)

    def checkpoint(self, n_epi, mean_Return, mode=None):
        """models_basic_lstm.py:375-391: same file names and dict keys."""
        if mode == "best":
            fname = self.tp["base_checkpoint_path"] + "best_model.tar"
            with open(self.tp["base_checkpoint_path"] + "/model_best_save_history.txt", "a") as f:
                f.write("BEST_timestep:" + str(n_epi) + ", avg det res:" + str(mean_Return) + "\n")
        elif mode == "last":
            fname = self.tp["base_checkpoint_path"] + "model_" + str(n_epi) + ".tar"
        else:
            return
        torch.save({"weights": self.state_dict(), "optimizer": self.optimizer.state_dict()}, fname)

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


class S2VKernelMixin:
    """The kernel-backed replacements for process_input / s2v / create_embeddings / pi / v.
    Mix in FRONT of the reference class to keep its make_batch / train_net / learn:

        class Model(S2VKernelMixin, reference.PPO_GNN_Single_LSTM): pass

    Needs only the attributes the reference constructor creates (theta*, lstm, config, emb_dim, T)."""

    # -- helpers ----------------------------------------------------------------------------
    @property
    def device(self):
        return self.theta6_pi.weight.device

    def embed_weights(self):
        return [t for n in ops.EMBED_NAMES for t in (getattr(self, n).weight, getattr(self, n).bias)]

    def head_weights(self, suffix):
        return [t for n in ops.HEAD_NAMES for t in (getattr(self, n + suffix).weight, getattr(self, n + suffix).bias)]

    def _layout(self, ei: torch.Tensor, num_nodes: int, copies: int) -> GraphLayout:
        """CSR of one graph, cached per edge_index tensor (a rollout shares one ei for all
        its steps and all of v', v, pi; models_basic_lstm.py:606-624).  For qnet='gat2' the
        self loops are stripped (GATv2Conv removes them and adds one per node; the kernels add it)."""
        if not hasattr(self, "_layouts"):
            self._layouts = {}
        gat = self.config["qnet"] == "gat2"
        key = (id(ei), ei._version, num_nodes, str(self.device), gat)
        hit = self._layouts.get(key)
        if hit is None:
            if len(self._layouts) >= 64:
                self._layouts.clear()
            dev_ei = ei.to(self.device)
            base = GraphLayout.replicated(strip_self_loops(dev_ei) if gat else dev_ei, num_nodes, 1, edge_map=gat)
            hit = (ei, base)                   # keep ei alive so id() stays unique
            self._layouts[key] = hit
        return hit[1].replicate(copies)

    # -- reference contract -----------------------------------------------------------------
    def process_input(self, nfm, ei, reachable, hidden):
        """models_basic_lstm.py:464-477."""
        dev = self.device
        if dev.type != "cuda":
            raise S2VError("simmodel_b200.PPO_GNN_Single_LSTM runs on CUDA only (no CPU fallback)")
        nfm = nfm.to(dev)
        reachable = reachable.to(dev)
        if nfm.dim() == 2:
            nfm = nfm[None, :, :]
            reachable = reachable[None, :]
        seq_len, num_nodes = nfm.shape[0], nfm.shape[1]
        if self.config["lstm_type"] == "FE":
            nfm, lstm_hidden = self.lstm(nfm, hidden)
        else:
            lstm_hidden = hidden
        return num_nodes, seq_len, nfm, ei, reachable, lstm_hidden

    def s2v(self, xv, Ws):
        """xv f32[S,N,F], Ws f32[N,N] shared 0/1 adjacency -> mu f32[S,N,p]
        (models_basic_lstm.py:479-505)."""
        dev = self.device
        Ws = Ws.to(dev)
        if not bool(((Ws == 0) | (Ws == 1)).all()):
            raise NotImplementedError("weighted / duplicate-edge adjacency: the reference graphs carry 0/1 weights only")
        ei = Ws.nonzero().t().contiguous()
        S, N = int(xv.shape[0]), int(xv.shape[1])
        layout = GraphLayout.replicated(ei, N, S)
        mu = ops.s2v_embed(layout, xv.to(dev).reshape(S * N, -1).contiguous(), self.embed_weights(), self.T)
        return mu.view(S, N, self.emb_dim)

    def create_embeddings(self, seq_len, nfm, ei, lstm_hidden):
        """models_basic_lstm.py:507-524 (s2v branch): returns (mu f32[S,N,p], hidden)."""
        mu, lstm_hidden, _ = self._create_embeddings(seq_len, nfm.to(self.device), ei, lstm_hidden)
        return mu, lstm_hidden

    def _create_embeddings(self, seq_len, nfm, ei, lstm_hidden):
        # CSR instead of to_dense_adj (models_basic_lstm.py:515)
        assert ei.dim() == 2
        num_nodes = int(nfm.shape[1])
        layout = self._layout(ei, num_nodes, int(seq_len))
        x = nfm.reshape(seq_len * num_nodes, -1).contiguous()
        if self.config["qnet"] == "gat2":          # Batch of S x Data(e, ei) -> self.gat(x, edge_index), :509-513
            mu = self.gat.forward_layout(x, layout)
        else:
            mu = ops.s2v_embed(layout, x, self.embed_weights(), self.T)
        mu = mu.view(seq_len, num_nodes, self.emb_dim)
        if self.config["lstm_type"] == "EMB":
            mu, lstm_hidden = self.lstm(mu, lstm_hidden)
        return mu, lstm_hidden, layout

    def pi(self, nfm, ei, reachable, hidden):
        """Masked policy: prob f32[S,N] (models_basic_lstm.py:526-542)."""
        num_nodes, seq_len, nfm, ei, reachable, lstm_hidden = self.process_input(nfm, ei, reachable, hidden)
        mu, lstm_hidden, layout = self._create_embeddings(seq_len, nfm, ei, lstm_hidden)
        # logit head + reachable-node mask + per-graph softmax: one launch (ops.s2v_head_softmax)
        prob = ops.s2v_head_softmax(layout, mu.reshape(seq_len * num_nodes, self.emb_dim), self.head_weights("_pi"), reachable)
        return prob.view(seq_len, num_nodes), lstm_hidden

    def v(self, nfm, ei, reachable, hidden):
        """State value f32[S,1] (models_basic_lstm.py:544-566)."""
        num_nodes, seq_len, nfm, ei, reachable, lstm_hidden = self.process_input(nfm, ei, reachable, hidden)
        mu, lstm_hidden, layout = self._create_embeddings(seq_len, nfm, ei, lstm_hidden)
        flat = mu.reshape(seq_len * num_nodes, self.emb_dim)
        if self.config["critic"] == "v":
            pool = ops.segment_pool(layout, flat, True)                       # mu.mean(dim=1)
            v = F.linear(pool, self.theta5_v.weight, self.theta5_v.bias)      # (S,1): B scalars, stays torch
        else:
            v, _ = ops.s2v_head_max(layout, flat, self.head_weights("_v"), reachable)    # Q head + mask + max: one launch
            v = v.unsqueeze(-1)
        return v, lstm_hidden

    # -- batched rollouts (caller-side batching of the reference's per-rollout loop) ---------------------
    def evaluate_rollouts(self, nfm, nfm_prime, ei, reachable, reachable_prime, first_hidden=None, second_hidden=None,
                          prime_is_shift=False):
        """What one PPO epoch needs from R rollouts that share a topology, in ONE pass over the GNN
        (models_basic_lstm.py:606-624 calls v(s'), v(s), pi(s) per rollout: three embedding passes each).

        nfm, nfm_prime f32[R,S,N,F]; reachable, reachable_prime bool[R,S,N]; hidden = (h, c) each [1, R*N, p]
        (rollout-major: rollout r owns rows r*N..(r+1)*N) or None.  prime_is_shift: the caller guarantees
        nfm_prime[:, :-1] == nfm[:, 1:] (consecutive steps of one episode), so only the last s' is embedded anew.
        Returns pi(s) f32[R,S,N], v(s) f32[R,S,1], v(s') f32[R,S,1] -- per graph-step the same arithmetic as pi()/v():
        the GNN embedding of a step does not depend on the batch around it, pi and v over s share embedding and
        LSTM output, and the LSTM (cuDNN, out of scope) runs with batch R*N instead of N."""
        dev = self.device
        if dev.type != "cuda":
            raise S2VError("simmodel_b200.PPO_GNN_Single_LSTM runs on CUDA only (no CPU fallback)")
        nfm, nfm_prime = nfm.to(dev), nfm_prime.to(dev)
        R, S, N, Fd = nfm.shape
        p = self.emb_dim
        if self.config["lstm_type"] == "FE":
            # the LSTM runs over the node FEATURES before the GNN (process_input, models_basic_lstm.py:473): s and s'
            # are two sequences with their own initial states, so the shifted-sequence shortcut does not apply; the
            # embedding backward returns dx and the LSTM trains through the kernels (s2v_embed_backward_dx)
            def run_fe(m, hidden):                                             # (seq S, batch R*N, F), rollout-major batch
                out, _ = self.lstm(m.permute(1, 0, 2, 3).reshape(S, R * N, Fd), hidden)
                return out.view(S, R, N, Fd).permute(1, 0, 2, 3)
            nfm, nfm_prime = run_fe(nfm, first_hidden), run_fe(nfm_prime, second_hidden)
            prime_is_shift = False
        extra = nfm_prime[:, -1:] if prime_is_shift else nfm_prime
        allx = torch.cat((nfm, extra), dim=1)                                  # [R, S+1 | 2S, N, F]
        G = int(allx.shape[1])
        layout = self._layout(ei, N, R * G)
        x = allx.reshape(R * G * N, Fd).contiguous()
        if self.config["qnet"] == "gat2":
            mu = self.gat.forward_layout(x, layout)
        else:
            mu = ops.s2v_embed(layout, x, self.embed_weights(), self.T)
        mu = mu.view(R, G, N, p)
        mu_s = mu[:, :S]
        mu_p = mu[:, 1:S + 1] if prime_is_shift else mu[:, S:]
        if self.config["lstm_type"] == "EMB":
            def run_lstm(m, hidden):                                           # (seq S, batch R*N, p), rollout-major batch
                out, _ = self.lstm(m.permute(1, 0, 2, 3).reshape(S, R * N, p), hidden)
                return out.view(S, R, N, p).permute(1, 0, 2, 3)
            mu_s, mu_p = run_lstm(mu_s, first_hidden), run_lstm(mu_p, second_hidden)
        lay_s = self._layout(ei, N, R * S)
        flat_s = mu_s.reshape(R * S * N, p).contiguous()
        flat_p = mu_p.reshape(R * S * N, p).contiguous()
        reach_s, reach_p = reachable.to(dev).reshape(R * S, N), reachable_prime.to(dev).reshape(R * S, N)
        prob = ops.s2v_head_softmax(lay_s, flat_s, self.head_weights("_pi"), reach_s).view(R, S, N)
        if self.config["critic"] == "v":
            v_s = F.linear(ops.segment_pool(lay_s, flat_s, True), self.theta5_v.weight, self.theta5_v.bias)
            v_p = F.linear(ops.segment_pool(lay_s, flat_p, True), self.theta5_v.weight, self.theta5_v.bias)
        else:
            v_s, _ = ops.s2v_head_max(lay_s, flat_s, self.head_weights("_v"), reach_s)
            v_p, _ = ops.s2v_head_max(lay_s, flat_p, self.head_weights("_v"), reach_p)
        return prob, v_s.reshape(R, S, 1), v_p.reshape(R, S, 1)

    @torch.no_grad()
    def greedy_action(self, nfm, ei, reachable, hidden):
        """probs.argmax() of the deterministic policy (rl_policy.py:533-536), on the GPU."""
        num_nodes, seq_len, nfm, ei, reachable, lstm_hidden = self.process_input(nfm, ei, reachable, hidden)
        mu, lstm_hidden, layout = self._create_embeddings(seq_len, nfm, ei, lstm_hidden)
        arg, _ = ops.s2v_head_argmax(layout, mu.reshape(seq_len * num_nodes, self.emb_dim), self.head_weights("_pi"), reachable)
        return arg, lstm_hidden


class PPO_GNN_Single_LSTM(S2VKernelMixin, _PPOModelBase):
    """Drop-in for models_basic_lstm.py:407 PPO_GNN_Single_LSTM with config['qnet'] in ('s2v', 'gat2')."""
