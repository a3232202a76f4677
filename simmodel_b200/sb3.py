"""Drop-in for the SB3 / MaskablePPO consumer of the s2v body.

Mirrors modules/ppo/models_sb3_s2v.py: ``Struc2VecExtractor`` (32-148: the s2v embedding over the PADDED
observation batch + serialisation), ``s2v_ACNetwork`` (150-238: actor / critic heads with SUM pooling over all
padded rows, critic = max over reachable nodes) and ``DeployablePPOPolicy`` (361-420: -1e20 masking, greedy action).
Same constructor arguments, parameter names and tensor shapes, so the state dicts of a trained SB3 policy load as
they are.  Conventions that differ from the DQN / PPO-LSTM consumers and are reproduced here:

  * padded rows are NOT dropped: they are isolated zero-feature nodes whose (non-zero) embedding enters the sum pool
    and gets logits of its own -- the layout therefore holds B graphs of exactly Npad nodes;
  * the mean pool shipped in the serialised features is over the real nodes only (scatter_mean over ``select``);
  * unreachable logits are set to -1e20 (not -inf) before Categorical / argmax.

The embedding, pooling, heads and masked max run in the sm_100a kernels (ops.s2v_embed / s2v_head / segment_pool /
masked_max / masked_argmax); torch only slices and concatenates.  CUDA only.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch.distributions.categorical import Categorical

from . import ops
from ._lib import S2VError
from .layout import GraphLayout


def _empty_ei(dev):
    return torch.zeros(2, 0, dtype=torch.int64, device=dev)


class _RangeLayouts:
    """B graphs of exactly Npad nodes, no edges: all the pooling / head / mask kernels need."""

    def __init__(self):
        self._d = {}

    def get(self, n_pad, B, dev) -> GraphLayout:
        key = (int(n_pad), str(dev))
        base = self._d.get(key)
        if base is None:
            base = self._d[key] = GraphLayout.replicated(_empty_ei(dev), int(n_pad), 1)
        return base.replicate(int(B))


class Struc2VecExtractor(nn.Module):
    """models_sb3_s2v.py:32-148.  ``observation_space``: a gym Dict space (``.spaces['W'].shape[0]`` = max nodes) or the
    max node count itself."""

    def __init__(self, observation_space, emb_dim, emb_iter_T, node_dim):
        super().__init__()
        max_num_nodes = observation_space if isinstance(observation_space, int) else observation_space.spaces["W"].shape[0]
        self._features_dim = max_num_nodes * (emb_dim + 1)
        self.norm_agg = True
        self.fdim = self._features_dim
        self.emb_dim, self.T, self.node_dim = emb_dim, emb_iter_T, node_dim
        self.theta1a = nn.Linear(node_dim, emb_dim, True)
        self.theta1b = nn.Linear(emb_dim, emb_dim, True)
        self.theta2 = nn.Linear(emb_dim, emb_dim, True)
        self.theta3 = nn.Linear(emb_dim, emb_dim, True)
        self.theta4 = nn.Linear(1, emb_dim, True)

    @property
    def features_dim(self):
        return self._features_dim

    def embed_weights(self):
        return [t for n in ops.EMBED_NAMES for t in (getattr(self, n).weight, getattr(self, n).bias)]

    def _padded_layout(self, Ws):
        B, n_pad = int(Ws.shape[0]), int(Ws.shape[1])
        if not bool(((Ws == 0) | (Ws == 1)).all()):
            raise NotImplementedError("weighted adjacency: the reference graphs carry 0/1 weights only")
        nz = Ws.nonzero()                                   # (b, i, j), row-major
        off = nz[:, 0] * n_pad
        ei = torch.stack((nz[:, 1] + off, nz[:, 2] + off))
        ptr = torch.arange(B + 1, device=Ws.device, dtype=torch.int64) * n_pad
        return GraphLayout.from_batch(ei, ptr, n_pad)

    def get_embeddings(self, xv, Ws):
        """xv f32[B,Npad,F], Ws f32[B,Npad,Npad] -> mu f32[B,Npad,p], padded rows included (:57-83)."""
        dev = self.theta2.weight.device
        if dev.type != "cuda":
            raise S2VError("simmodel_b200.Struc2VecExtractor runs on CUDA only (no CPU fallback)")
        xv, Ws = xv.to(dev, torch.float32), Ws.to(dev)
        B, n_pad = int(xv.shape[0]), int(xv.shape[1])
        layout = self._padded_layout(Ws)
        mu = ops.s2v_embed(layout, xv.reshape(B * n_pad, -1).contiguous(), self.embed_weights(), self.T)
        return mu.view(B, n_pad, self.emb_dim)

    def forward(self, observations):
        """-> mu_serialized f32[B, Npad+1, p+1] (:85-148)."""
        dev = self.theta2.weight.device
        nfm, W = observations["nfm"], observations["W"]
        num_nodes = observations["num_nodes"].to(dev)
        n_pad, B = int(W.shape[1]), int(W.shape[0])
        mu = self.get_embeddings(nfm, W)
        sizes = num_nodes.reshape(-1).to(torch.int64)
        ptr = torch.zeros(B + 1, dtype=torch.int64, device=dev)
        ptr[1:] = torch.cumsum(sizes, 0)
        batch = torch.repeat_interleave(torch.arange(B, device=dev), sizes)
        select = batch * n_pad + (torch.arange(int(ptr[-1]), device=dev) - ptr[batch])     # the real rows (:104-111)
        mu_raw = mu.reshape(-1, self.emb_dim)[select]
        real = GraphLayout.from_batch(_empty_ei(dev), ptr, n_pad, batch)
        mu_meanpool = ops.segment_pool(real, mu_raw, True)                                   # scatter_mean (:114)
        ser1 = torch.cat((mu, observations["reachable_nodes"].to(dev, mu.dtype)), dim=2)
        ser2 = torch.cat((mu_meanpool, num_nodes.to(mu.dtype)), dim=1)[:, None, :]
        return torch.cat((ser1, ser2), dim=1)


class s2v_ACNetwork(nn.Module):
    """models_sb3_s2v.py:150-238."""

    def __init__(self, feature_dim, last_layer_dim_pi=64, last_layer_dim_vf=64, emb_dim=0):
        super().__init__()
        self.latent_dim_pi, self.latent_dim_vf = last_layer_dim_pi, last_layer_dim_vf
        self.feature_dim, self.emb_dim = feature_dim, emb_dim
        self.theta5_pi = nn.Linear(2 * emb_dim, 1, True)
        self.theta6_pi = nn.Linear(emb_dim, emb_dim, True)
        self.theta7_pi = nn.Linear(emb_dim, emb_dim, True)
        self.theta5_v = nn.Linear(2 * emb_dim, 1, True)
        self.theta6_v = nn.Linear(emb_dim, emb_dim, True)
        self.theta7_v = nn.Linear(emb_dim, emb_dim, True)
        self._ranges = _RangeLayouts()

    def head_weights(self, suffix):
        return [t for n in ops.HEAD_NAMES for t in (getattr(self, n + suffix).weight, getattr(self, n + suffix).bias)]

    def _deserialize(self, features):
        s0, s1, s2 = features.shape
        a, b = torch.split(features, [s2 - 1, 1], dim=2)
        mu, mu_mp = torch.split(a, [s1 - 1, 1], dim=1)
        reachable_nodes, num_nodes = torch.split(b, [s1 - 1, 1], dim=1)
        return mu, mu_mp.squeeze(), reachable_nodes, num_nodes.squeeze(2), s1 - 1, s2 - 1, s0

    def _flat(self, mu):
        if not mu.is_cuda:
            raise S2VError("simmodel_b200.s2v_ACNetwork runs on CUDA only (no CPU fallback)")
        B, n_pad, p = mu.shape
        return self._ranges.get(n_pad, B, mu.device), mu.reshape(B * n_pad, p).contiguous()

    def forward(self, features):
        mu, _, reachable_nodes, _, n_pad, _, _ = self._deserialize(features)
        return self.forward_actor(features, mu, n_pad), self.forward_critic(features, mu, reachable_nodes, n_pad)

    def forward_actor(self, features, mu=None, num_nodes_padded=None):
        if mu is None:
            mu = self._deserialize(features)[0]
        layout, flat = self._flat(mu)
        logits = ops.s2v_head(layout, flat, self.head_weights("_pi"), False)       # sum pool over ALL padded rows (:216)
        return logits.view(mu.shape[0], mu.shape[1], 1)

    def forward_critic(self, features, mu=None, reachable_nodes=None, num_nodes_padded=None):
        if mu is None:
            mu, _, reachable_nodes, _, _, _, _ = self._deserialize(features)
        layout, flat = self._flat(mu)
        qvals = ops.s2v_head(layout, flat, self.head_weights("_v"), False)
        v, _ = ops.masked_max(layout, qvals, reachable_nodes.reshape(mu.shape[0], mu.shape[1]) != 0)   # (:234-236)
        return v.unsqueeze(-1)


class DeployablePPOPolicy(nn.Module):
    """models_sb3_s2v.py:361-420: node-count-invariant deployment of a trained SB3 policy."""

    def __init__(self, env, trained_policy, device="cuda"):
        super().__init__()
        self.device = device
        self.struc2vec_extractor = Struc2VecExtractor(env.observation_space, 64, 5, env.F).to(device)
        self.struc2vec_extractor.load_state_dict(trained_policy.features_extractor.state_dict())
        self.s2vACnet = s2v_ACNetwork(64, 1, 1, 64).to(device)
        self.s2vACnet.load_state_dict(trained_policy.mlp_extractor.state_dict())
        self.pnet = nn.Linear(1, 1, True).to(device)
        self.pnet.load_state_dict(trained_policy.action_net.state_dict())
        self.vnet = nn.Linear(1, 1, True).to(device)
        self.vnet.load_state_dict(trained_policy.value_net.state_dict())

    def forward(self, obs):
        a, b = self.s2vACnet(self.struc2vec_extractor(obs))
        return self.pnet(a), self.vnet(b)                  # (B,Npad,1) logits, (B,1) value: 1x1 affine maps stay torch

    def predict(self, obs, deterministic=True, action_masks=None):
        raw_logits, _ = self.forward(obs)
        assert deterministic
        flat = raw_logits.detach().reshape(-1)
        m = torch.as_tensor(np.asarray(action_masks), dtype=torch.bool, device=flat.device).reshape(-1)
        layout = self.s2vACnet._ranges.get(flat.numel(), 1, flat.device)
        arg, _ = ops.masked_argmax(layout, flat, m)        # == argmax(where(m, logits, -1e20)): first maximum
        action = int(arg.item())
        return np.asarray(max(action, 0)), None            # nothing reachable: every entry is -1e20, argmax is 0

    def get_distribution(self, obs):
        raw_logits, _ = self.forward(obs)
        m = obs["reachable_nodes"].to(raw_logits.device).to(torch.bool)
        huge_neg = torch.tensor(-1e20, dtype=torch.float32, device=raw_logits.device)
        return Categorical(logits=torch.where(m.squeeze(), raw_logits.squeeze(-1), huge_neg))

    def predict_values(self, obs):
        return self.forward(obs)[1]
