"""Generate tests/golden/gat_*.npz for the GATv2 branch (TEST INFRASTRUCTURE ONLY; build container only:
python -m oracle.gen_golden_gat).

The reference's OWN classes run unmodified -- ``QNet_GATv2`` (modules/gnn/comb_opt.py:156-182, forward 84-135) and
``PPO_GNN_Single_LSTM`` with config['qnet'] == 'gat2' (modules/ppo/models_basic_lstm.py:407-566) -- with the absent
third-party PyG classes (GATv2Conv, BasicGNN, Data, Batch) replaced by the restatements of oracle/gat_ref.py.  So the
fixtures pin the reference's call sites, batching, pooling, heads and masks; the conv arithmetic itself is the
published PyG algorithm restated (UNPINNED against real PyG, which cannot be installed here -- see gat_ref.py).
Also extracts the reference's shipped, trained gat2 checkpoint into tests/golden/gat2_ppo_checkpoint.npz and runs it.
"""
from __future__ import annotations

import contextlib
import glob
import io
import os
import types

import numpy as np
import torch

from oracle import gat_ref as G
from oracle.gen_golden import _np, _random_digraph, _save, _scale_of
from oracle.ref_import import REFERENCE_ROOT, load_reference
from simmodel_b200 import graphs


def _batch(xs, eis):
    return G.batch_from_data_list([G.Data(x, ei) for x, ei in zip(xs, eis)])


def dqn_case(co, name, xs, eis, p, norm_agg, seed, scale=1.0):
    torch.manual_seed(seed)
    config = {"emb_dim": p, "emb_iter_T": 5, "norm_agg": norm_agg, "node_dim": graphs.NODE_DIM}
    with contextlib.redirect_stdout(io.StringIO()):
        net = co.QNet_GATv2(config)
    with torch.no_grad():                       # PyG zero-initialises the conv bias: make it count
        for k, prm in net.named_parameters():
            if k.endswith(".bias") and "convs" in k and "lin_" not in k:
                prm.copy_(torch.randn(prm.shape) * 0.1)
            if scale != 1.0:
                prm.mul_(scale)
    P = {k: v.detach().clone() for k, v in net.state_dict().items()}
    batch = _batch(xs, eis)
    sizes = [int(x.shape[0]) for x in xs]
    B = len(xs)
    gen = torch.Generator().manual_seed(seed + 1)
    q = net(None, None, batch)
    actions = [int(torch.randint(0, s, (1,), generator=gen)) for s in sizes]
    targets = torch.randn(B, generator=gen)
    sel = torch.tensor(actions, dtype=torch.int64) + batch.ptr[:-1]
    loss = torch.nn.SmoothL1Loss()(q.reshape(-1)[sel], targets)
    net.zero_grad()
    loss.backward()
    grads = {k: prm.grad.detach().clone() for k, prm in net.named_parameters()}
    # functional restatement == the reference driving the injected conv, bit for bit
    qo = G.qnet_gat_forward(P, batch.x, batch.edge_index, batch.ptr, 5, norm_agg)
    assert torch.equal(qo.reshape(-1), q.detach().reshape(-1)), "functional GAT restatement differs from the reference run"
    # fp64 adjudication
    P64 = {k: v.double().requires_grad_(True) for k, v in P.items()}
    q64 = G.qnet_gat_forward(P64, batch.x.double(), batch.edge_index, batch.ptr, 5, norm_agg).reshape(-1)
    torch.nn.SmoothL1Loss()(q64[sel], targets.double()).backward()
    g64 = {k: P64[k].grad.detach() for k in grads}
    ref_err = float((q.detach().reshape(-1).double() - q64.detach()).abs().max()) / _scale_of(q64)
    ref_gerr = max(float((grads[k].double() - g64[k]).abs().max()) / _scale_of(g64[k]) for k in grads)
    if os.environ.get("GAT_GOLDEN_VERBOSE"):
        for k in grads:
            print("      %-28s scale %.3g err %.1e" % (k, _scale_of(g64[k]), float((grads[k].double() - g64[k]).abs().max()) / _scale_of(g64[k])))
    print("  %-28s q scale %.3g | reference's own fp32 error vs fp64: q %.1e grads %.1e" % (name, _scale_of(q), ref_err, ref_gerr))
    meta = dict(kind="gat_dqn", sizes=sizes, emb_dim=p, norm_agg=norm_agg, seed=seed)
    _save(name, meta, P={k: _np(v) for k, v in P.items()},
          inp=dict(x=_np(batch.x), edge_index=_np(batch.edge_index).astype(np.int32), ptr=_np(batch.ptr).astype(np.int32),
                   actions=np.array(actions, dtype=np.int64), targets=_np(targets)),
          out=dict(q=_np(q).reshape(-1), loss=_np(loss), q_f64=_np(q64)),
          grad={k: _np(v) for k, v in grads.items()}, grad64={k: _np(v) for k, v in g64.items()})


def ppo_case(mb, name, topo, S, p, T, lstm_type, critic, seed, state=None):
    torch.manual_seed(seed)
    config = {"lstm_type": lstm_type, "qnet": "gat2", "critic": critic, "emb_dim": p, "emb_iterT": T,
              "gat_concat": True, "gat_heads": 4, "gat_share_weights": False}       # ppo_custom.py:254-257
    hp = types.SimpleNamespace(node_dim=graphs.NODE_DIM, emb_dim=p, parallel_rollouts=4, batch_size=2, learning_rate=1e-3)
    tp = {"base_checkpoint_path": "/tmp/simmodel_golden_ckpt/"}
    with contextlib.redirect_stdout(io.StringIO()):
        model = mb.PPO_GNN_Single_LSTM(config, hp, tp)
    if state is not None:
        model.load_state_dict(state, strict=True)                                   # the shipped checkpoint
    else:
        with torch.no_grad():
            for k, prm in model.named_parameters():
                if k.endswith(".bias") and "convs" in k and "lin_" not in k:
                    prm.copy_(torch.randn(prm.shape) * 0.1)
    P = {k: v.detach().clone() for k, v in model.state_dict().items()}
    gen = torch.Generator().manual_seed(seed + 1)
    N = topo.num_nodes
    nfm = graphs.synth_nfm(topo, S, gen)
    ei = topo.edge_index
    reachable = torch.zeros(S, N, dtype=torch.bool)
    for t in range(S):
        cur = int(nfm[t, :, 1].argmax())
        reachable[t, topo.neighbors(cur) or [cur]] = True                           # environments.py:421-425
    hidden = None
    if lstm_type == "EMB":
        hidden = (torch.randn(1, N, p, generator=gen) * 0.1, torch.randn(1, N, p, generator=gen) * 0.1)
    prob, _ = model.pi(nfm, ei, reachable, hidden)
    v, _ = model.v(nfm, ei, reachable, hidden)
    wprob = torch.randn(S, N, generator=gen)
    wv = torch.randn(S, 1, generator=gen)
    loss = (prob * wprob).sum() + (v * wv).sum()
    model.zero_grad()
    loss.backward()
    grads = {k: (prm.grad.detach().clone() if prm.grad is not None else torch.zeros_like(prm)) for k, prm in model.named_parameters()}
    mu, _ = model.create_embeddings(S, nfm, ei, hidden)
    # functional restatement of the embedding == the reference run, bit for bit
    bt = _batch(list(nfm), [ei] * S)
    mu_gat = G.gat_forward(P, bt.x, bt.edge_index, T).reshape(S, N, p)
    if lstm_type != "EMB":
        assert torch.equal(mu_gat, mu.detach()), "functional GAT restatement differs from the reference run"
    # fp64 adjudication: the whole pi / v arithmetic restated functionally in double (LSTM included)
    P64 = {k: v_.double().requires_grad_(True) for k, v_ in P.items()}
    mu64 = G.gat_forward(P64, bt.x.double(), bt.edge_index, T).reshape(S, N, p)
    own = float((mu_gat.double() - mu64.detach()).abs().max()) / _scale_of(mu64)
    emb64 = mu64
    if lstm_type == "EMB":
        emb64 = torch._VF.lstm(mu64, (hidden[0].double(), hidden[1].double()),
                               [P64["lstm.weight_ih_l0"], P64["lstm.weight_hh_l0"], P64["lstm.bias_ih_l0"],
                                P64["lstm.bias_hh_l0"]], True, 1, 0.0, False, False, False)[0]
    prob64, v64 = G.ppo_heads(P64, emb64, reachable, critic)
    ((prob64 * wprob.double()).sum() + (v64 * wv.double()).sum()).backward()
    g64 = {k: (P64[k].grad.detach() if P64[k].grad is not None else torch.zeros_like(P64[k])) for k in grads}
    gown = max(float((grads[k].double() - g64[k]).abs().max()) / _scale_of(g64[k]) for k in grads if _scale_of(g64[k]) > 1e-12)
    print("  %-28s prob max %.3g v scale %.3g | reference's own fp32 error vs fp64: embedding %.1e prob %.1e v %.1e grads %.1e" % (
        name, float(prob.max()), _scale_of(v), own, float((prob.detach().double() - prob64.detach()).abs().max()),
        float((v.detach().double() - v64.detach()).abs().max()) / _scale_of(v64), gown))
    meta = dict(kind="gat_ppo", S=S, N=N, emb_dim=p, T=T, lstm_type=lstm_type, critic=critic, seed=seed, topo=topo.name)
    inp = dict(nfm=_np(nfm), edge_index=_np(ei).astype(np.int32), reachable=_np(reachable), wprob=_np(wprob), wv=_np(wv))
    if hidden is not None:
        inp["h0"], inp["c0"] = _np(hidden[0]), _np(hidden[1])
    groups = dict(inp=inp,
                  out=dict(prob=_np(prob), v=_np(v), mu=_np(mu), mu_gat=_np(mu_gat), greedy=_np(prob.detach().argmax(dim=1)),
                           prob_f64=_np(prob64), v_f64=_np(v64)),
                  grad={k: _np(g) for k, g in grads.items()},
                  grad64={k: _np(g).astype(np.float32 if state is not None else np.float64) for k, g in g64.items()})
    if state is None:
        groups["P"] = {k: _np(v_) for k, v_ in P.items()}
    _save(name, meta, **groups)


def star_digraph(n, gen):
    """random digraph + node 0 connected with everyone in both directions (in-degree n-1 > 8: looped kernel path),
    with a few self loops (GATv2Conv removes them)."""
    ei = _random_digraph(n, 2.0, gen)
    hub = torch.arange(1, n)
    extra = torch.cat((torch.stack((hub, torch.zeros_like(hub))), torch.stack((torch.zeros_like(hub), hub)),
                       torch.tensor([[0, 3], [0, 3]])), dim=1)
    ei = torch.cat((ei, extra), dim=1)
    key = ei[0] * n + ei[1]
    uniq = torch.unique(key)                    # sorted: row-major like W.nonzero()
    return torch.stack((uniq // n, uniq % n))


def main():
    co, mb = load_reference()
    gen = torch.Generator().manual_seed(4321)
    print("GATv2 DQN cases (QNet_GATv2)")
    m5 = graphs.load_topology("Manhattan5")
    x5 = graphs.synth_nfm(m5, 4, gen)
    dqn_case(co, "gat_dqn_manhattan5_b4", list(x5), [m5.edge_index] * 4, 64, True, seed=20)
    sizes = [5, 14, 1, 12, 7]
    eis = [star_digraph(s, gen) if s >= 12 else _random_digraph(s, 2.5, gen) for s in sizes]
    xs = [torch.randn(s, graphs.NODE_DIM, generator=gen).abs() for s in sizes]
    dqn_case(co, "gat_dqn_varsize_star_p16_sum", xs, eis, 16, False, seed=21)
    dqn_case(co, "gat_dqn_varsize_star_p64", xs, eis, 64, True, seed=22)
    ams = graphs.load_topology("NWB_AMS")
    xa = graphs.synth_nfm(ams, 2, gen)
    dqn_case(co, "gat_dqn_ams_b2", list(xa), [ams.edge_index] * 2, 64, True, seed=23)

    print("GATv2 PPO cases (PPO_GNN_Single_LSTM, qnet='gat2')")
    mem = graphs.load_topology("MemTask")
    ppo_case(mb, "gat_ppo_memtask_p24_none_q", mem, 5, 24, 5, "None", "q", seed=30)
    ppo_case(mb, "gat_ppo_manhattan5_p64_emb_q", m5, 6, 64, 5, "EMB", "q", seed=31)
    ppo_case(mb, "gat_ppo_memtask_p64_none_v_T3", mem, 5, 64, 3, "None", "v", seed=32)
    # the reference's shipped trained checkpoint (gat2-q, emb64, itT5, lstm EMB, NWB_AMS_mixed_obs)
    f = glob.glob(os.path.join(REFERENCE_ROOT, "results", "**", "gat2-q", "**", "best_model.tar"), recursive=True)[0]
    state = torch.load(f, map_location="cpu", weights_only=False)["weights"]
    out = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "gat2_ppo_checkpoint.npz")
    np.savez_compressed(out, **{k: _np(v) for k, v in state.items()})
    print("  wrote gat2_ppo_checkpoint (%d KB) from %s" % (os.path.getsize(out) // 1024, os.path.relpath(f, REFERENCE_ROOT)))
    ppo_case(mb, "gat_ppo_ams_s3_checkpoint", ams, 3, 64, 5, "EMB", "q", seed=33, state=state)


if __name__ == "__main__":
    main()
