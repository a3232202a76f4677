"""Golden vectors for lstm_type = 'FE' (an LSTM over the node features BEFORE the GNN, models_basic_lstm.py:442-444,
473): the unmodified reference PPO_GNN_Single_LSTM run on CPU.  TEST INFRASTRUCTURE ONLY; build container only:
    python -m oracle.gen_golden_fe
The FE configuration is the one path on which the embedding needs a gradient w.r.t. its INPUT features (the LSTM's
parameters train through the GNN), so these cases pin dx of the embedding backward (s2v_embed_backward_dx)."""
from __future__ import annotations

import contextlib
import io
import types

import numpy as np
import torch

from oracle import s2v_ref as R
from oracle.gen_golden import _np, _save, _scale_of
from oracle.ref_import import load_reference
from simmodel_b200 import graphs


def _lstm64(P64, x, h):
    return torch._VF.lstm(x, h, [P64["lstm.weight_ih_l0"], P64["lstm.weight_hh_l0"], P64["lstm.bias_ih_l0"],
                                 P64["lstm.bias_hh_l0"]], True, 1, 0.0, False, False, False)[0]


def fe_case(mb, name, topo, S, p, T, critic, seed):
    torch.manual_seed(seed)
    Fd = graphs.NODE_DIM
    config = {"lstm_type": "FE", "qnet": "s2v", "critic": critic, "emb_dim": p, "emb_iterT": T}
    hp = types.SimpleNamespace(node_dim=Fd, emb_dim=p, parallel_rollouts=4, batch_size=2, learning_rate=1e-3)
    with contextlib.redirect_stdout(io.StringIO()):
        model = mb.PPO_GNN_Single_LSTM(config, hp, {"base_checkpoint_path": "/tmp/simmodel_golden_ckpt/"})
    P = {k: v.detach().clone() for k, v in model.state_dict().items()}
    gen = torch.Generator().manual_seed(seed + 1)
    N = topo.num_nodes
    nfm = graphs.synth_nfm(topo, S, gen)
    ei = topo.edge_index
    reachable = torch.zeros(S, N, dtype=torch.bool)
    for t in range(S):
        cur = int(nfm[t, :, 1].argmax())
        reachable[t, topo.neighbors(cur) or [cur]] = True
    hidden = (torch.randn(1, N, Fd, generator=gen) * 0.1, torch.randn(1, N, Fd, generator=gen) * 0.1)
    prob, _ = model.pi(nfm, ei, reachable, hidden)
    v, _ = model.v(nfm, ei, reachable, hidden)
    wprob = torch.randn(S, N, generator=gen)
    wv = torch.randn(S, 1, generator=gen)
    model.zero_grad()
    ((prob * wprob).sum() + (v * wv).sum()).backward()
    grads = {k: (prm.grad.detach().clone() if prm.grad is not None else torch.zeros_like(prm))
             for k, prm in model.named_parameters()}
    # the dense restatement on the LSTM's output reproduces the reference bit for bit
    x_fe = model.lstm(nfm, hidden)[0].detach()        # grad mode on: the same CPU LSTM path as inside model.pi / model.v
    prob_o, _ = R.ppo_pi_dense(P, x_fe, ei, reachable, T, None, None)
    v_o, _ = R.ppo_v_dense(P, x_fe, ei, reachable, T, critic, None, None)
    assert torch.equal(prob_o, prob.detach()) and torch.equal(v_o, v.detach()), "dense PPO restatement differs"
    # fp64 adjudication of the same arithmetic, LSTM included
    P64 = {k: v_.double().requires_grad_(True) for k, v_ in P.items()}
    x64 = _lstm64(P64, nfm.double(), (hidden[0].double(), hidden[1].double()))
    prob64, _ = R.ppo_pi_dense(P64, x64, ei, reachable, T, None, None)
    v64, _ = R.ppo_v_dense(P64, x64, ei, reachable, T, critic, None, None)
    ((prob64 * wprob.double()).sum() + (v64 * wv.double()).sum()).backward()
    g64 = {k: (P64[k].grad.detach() if P64[k].grad is not None else torch.zeros_like(P64[k])) for k in grads}
    own = max(float((grads[k].double() - g64[k]).abs().max()) / _scale_of(g64[k]) for k in grads if _scale_of(g64[k]) > 1e-6)
    print("  %-26s reference's own fp32 error vs fp64: grads %.1e; lstm grad scale %.2e" % (
        name, own, _scale_of(g64["lstm.weight_ih_l0"])))
    meta = dict(kind="ppofe", S=S, N=N, emb_dim=p, T=T, lstm_type="FE", critic=critic, seed=seed, topo=topo.name)
    _save(name, meta, P={k: _np(v_) for k, v_ in P.items()},
          inp=dict(nfm=_np(nfm), edge_index=_np(ei).astype(np.int32), reachable=_np(reachable), wprob=_np(wprob), wv=_np(wv),
                   h0=_np(hidden[0]), c0=_np(hidden[1])),
          out=dict(prob=_np(prob), v=_np(v), prob_f64=_np(prob64), v_f64=_np(v64), greedy=_np(prob.detach().argmax(dim=1))),
          grad={k: _np(g) for k, g in grads.items()}, grad64={k: _np(g) for k, g in g64.items()})


def main():
    _, mb = load_reference()
    fe_case(mb, "ppofe_memtask_p24_q", graphs.load_topology("MemTask"), 5, 24, 5, "q", seed=20)
    fe_case(mb, "ppofe_manhattan5_p64_q", graphs.load_topology("Manhattan5"), 4, 64, 5, "q", seed=21)
    fe_case(mb, "ppofe_ams_s5_p64_v", graphs.load_topology("NWB_AMS"), 5, 64, 5, "v", seed=22)


if __name__ == "__main__":
    main()
