"""Generate tests/golden/*.npz by running the UNMODIFIED reference classes
(TEST INFRASTRUCTURE ONLY; build container only:  python -m oracle.gen_golden).

For each case the real ``QNet`` (modules/gnn/comb_opt.py:184-287), ``QFunction``
(289-366) or ``PPO_GNN_Single_LSTM`` (modules/ppo/models_basic_lstm.py:407-566)
is constructed under a fixed seed and run on seeded inputs in fp32 on CPU.  The
parameters, inputs, outputs and parameter gradients go into one npz per case.
While generating, the oracle restatement (oracle/s2v_ref.py) is checked against
the reference output: dense form bit-for-bit, sparse form to 2e-6 of scale.
"""
from __future__ import annotations

import json
import os
import types

import numpy as np
import torch
import torch.nn.functional as F

from oracle import s2v_ref as R
from oracle.ref_import import load_reference
from simmodel_b200 import graphs

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def _dense_W(ei, n_pad):
    W = torch.zeros(n_pad, n_pad, dtype=torch.float32)
    W[ei[0], ei[1]] = 1.0
    return W


def _save(name, meta, **groups):
    flat = {"meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)}
    for g, d in groups.items():
        for k, v in d.items():
            flat[g + "/" + k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **flat)
    print("  wrote %s (%d KB)" % (name, os.path.getsize(path) // 1024))


def _scale_of(t):
    return float(t.detach().abs().max()) + 1e-30


def _random_digraph(n, avg_deg, gen):
    m = torch.rand(n, n, generator=gen) < (avg_deg / max(n, 1))
    return m.nonzero().t().contiguous()


def dqn_case(co, name, sizes, eis, n_pad, p, T, norm_agg, seed, scale=1.0, xs=None, with_update=True):
    """sizes: real node count per graph; eis: per-graph local edge_index; n_pad: padded size."""
    torch.manual_seed(seed)
    config = {"emb_dim": p, "emb_iter_T": T, "norm_agg": norm_agg, "node_dim": graphs.NODE_DIM}
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        net = co.QNet(config)
    if scale != 1.0:
        with torch.no_grad():
            for prm in net.parameters():
                prm.mul_(scale)
    P = {k: v.detach().clone() for k, v in net.state_dict().items()}
    B = len(sizes)
    gen = torch.Generator().manual_seed(seed + 1)
    xv = torch.zeros(B, n_pad, graphs.NODE_DIM)
    Ws = torch.zeros(B, n_pad, n_pad)
    for b in range(B):
        xv[b, :sizes[b]] = xs[b] if xs is not None else torch.randn(sizes[b], graphs.NODE_DIM, generator=gen).abs()
        Ws[b] = _dense_W(eis[b], n_pad)
    batch = R.batch_from_sizes(sizes, eis)
    # --- the reference forward, the way batch_update drives it (comb_opt.py:349-362) ---
    q = net(xv, Ws, batch)
    actions = [int(torch.randint(0, s, (1,), generator=gen)) for s in sizes]
    targets = torch.randn(B, generator=gen).tolist()
    sel = torch.tensor(actions, dtype=torch.int64) + batch.ptr[:-1]
    loss = torch.nn.SmoothL1Loss()(q.reshape(-1)[sel], torch.tensor(targets))
    net.zero_grad()
    loss.backward()
    grads = {k: prm.grad.detach().clone() for k, prm in net.named_parameters()}
    # --- greedy action through the reference's own QFunction (comb_opt.py:309-331), B=1 unpadded ---
    qf = co.QFunction(net, None, None)
    n0 = sizes[0]
    reach0 = sorted(set(eis[0][1][eis[0][0] == actions[0]].tolist())) or [0]
    pyg0 = types.SimpleNamespace(num_nodes=n0)
    a_idx, a_node, a_val = qf.get_best_action(xv[0, :n0], Ws[0, :n0, :n0], pyg0, reach0)
    # --- oracle checks ---
    qd = R.qnet_forward_dense(P, xv, Ws, batch.ptr, batch.batch, T, norm_agg)
    assert torch.equal(qd.reshape(-1), q.detach().reshape(-1)), "dense restatement differs from reference"
    x_flat = torch.cat([xv[b, :sizes[b]] for b in range(B)])
    qs = R.qnet_forward_sparse(P, x_flat, batch.edge_index, batch.ptr, n_pad, T, norm_agg)
    err = float((qs - q.detach().reshape(-1)).abs().max()) / _scale_of(q)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    qs2 = R.qnet_forward_sparse(Pg, x_flat, batch.edge_index, batch.ptr, n_pad, T, norm_agg)
    R.dqn_loss(qs2, actions, batch.ptr, targets).backward()
    gerr = max(float((Pg[k].grad - grads[k]).abs().max()) / _scale_of(grads[k]) for k in grads)
    # fp64 adjudication: the reference's own fp32 rounding error vs the same arithmetic in fp64
    P64 = {k: v.double().requires_grad_(True) for k, v in P.items()}
    q64 = R.qnet_forward_dense(P64, xv.double(), Ws.double(), batch.ptr, batch.batch, T, norm_agg).reshape(-1)
    R.dqn_loss(q64, actions, batch.ptr, targets).backward()
    g64 = {k: P64[k].grad.detach() for k in grads}
    ref_err = float((q.detach().reshape(-1).double() - q64.detach()).abs().max()) / _scale_of(q)
    sp_err = float((qs.double() - q64.detach()).abs().max()) / _scale_of(q)
    ref_gerr = max(float((grads[k].double() - g64[k]).abs().max()) / _scale_of(g64[k]) for k in grads)
    sp_gerr = max(float((Pg[k].grad.double() - g64[k]).abs().max()) / _scale_of(g64[k]) for k in grads)
    assert sp_err < max(2e-6, 2 * ref_err), ("sparse restatement vs fp64", sp_err, ref_err)
    assert sp_gerr < max(2e-6, 2 * ref_gerr), ("sparse restatement grads vs fp64", sp_gerr, ref_gerr)
    print("  %-30s q scale %.3g | vs ref: q %.1e grads %.1e | vs fp64: ref q %.1e g %.1e, sparse q %.1e g %.1e"
          % (name, _scale_of(q), err, gerr, ref_err, ref_gerr, sp_err, sp_gerr))
    meta = dict(kind="dqn", sizes=sizes, n_pad=n_pad, emb_dim=p, T=T, norm_agg=norm_agg, seed=seed,
                greedy=dict(reachable=reach0, action=int(a_idx), node=int(a_node)))
    _save(name, meta,
          P={k: _np(v) for k, v in P.items()},
          inp=dict(x=_np(x_flat), edge_index=_np(batch.edge_index).astype(np.int32), ptr=_np(batch.ptr).astype(np.int32),
                   actions=np.array(actions, dtype=np.int64), targets=np.array(targets, dtype=np.float32)),
          out=dict(q=_np(q).reshape(-1), loss=_np(loss), greedy_value=np.float32(a_val), q_f64=_np(q64)),
          grad={k: _np(v) for k, v in grads.items()},
          grad64={k: _np(v) for k, v in g64.items()})


def ppo_case(mb, name, topo, S, p, T, lstm_type, critic, seed, obs_mask_freq=None):
    torch.manual_seed(seed)
    config = {"lstm_type": lstm_type, "qnet": "s2v", "critic": critic, "emb_dim": p, "emb_iterT": T}
    hp = types.SimpleNamespace(node_dim=graphs.NODE_DIM, emb_dim=p, parallel_rollouts=4, batch_size=2,
                               learning_rate=1e-3)
    tp = {"base_checkpoint_path": "/tmp/simmodel_golden_ckpt/"}
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        model = mb.PPO_GNN_Single_LSTM(config, hp, tp)
    P = {k: v.detach().clone() for k, v in model.state_dict().items()}
    gen = torch.Generator().manual_seed(seed + 1)
    N = topo.num_nodes
    nfm = graphs.synth_nfm(topo, S, gen)
    if obs_mask_freq:                       # ppo_wrappers.py:366-373: unit columns hidden off-frequency
        for t in range(S):
            if t % obs_mask_freq != 0:
                nfm[t, :, 5:7] = 0
    ei = topo.edge_index
    reachable = torch.zeros(S, N, dtype=torch.bool)
    for t in range(S):
        cur = int(nfm[t, :, 1].argmax())
        nb = topo.neighbors(cur) or [cur]
        reachable[t, nb] = True             # environments.py:421-425
    hidden = None
    if lstm_type == "EMB":
        hidden = (torch.randn(1, N, p, generator=gen) * 0.1, torch.randn(1, N, p, generator=gen) * 0.1)
    prob, h_pi = model.pi(nfm, ei, reachable, hidden)
    v, _ = model.v(nfm, ei, reachable, hidden)
    wprob = torch.randn(S, N, generator=gen)
    wv = torch.randn(S, 1, generator=gen)
    loss = (prob * wprob).sum() + (v * wv).sum()
    model.zero_grad()
    loss.backward()
    grads = {k: (prm.grad.detach().clone() if prm.grad is not None else torch.zeros_like(prm))
             for k, prm in model.named_parameters()}
    mu, _ = model.create_embeddings(S, nfm, ei, None if lstm_type != "EMB" else hidden)
    mu_s2v = model.s2v(nfm, R.to_dense_adj(ei)[0])
    # oracle checks (dense restatement bit-for-bit)
    lstm = model.lstm if lstm_type == "EMB" else None
    prob_o, _ = R.ppo_pi_dense(P, nfm, ei, reachable, T, lstm, hidden)
    v_o, _ = R.ppo_v_dense(P, nfm, ei, reachable, T, critic, lstm, hidden)
    assert torch.equal(prob_o, prob.detach()) and torch.equal(v_o, v.detach()), "dense PPO restatement differs"
    greedy = prob.detach().argmax(dim=1)    # rl_policy.py:536
    # fp64 adjudication: the same arithmetic (oracle dense restatement, LSTM included) in double
    P64 = {k: v_.double().requires_grad_(True) for k, v_ in P.items()}
    lstm64, hidden64 = None, None
    if lstm_type == "EMB":
        lstm64 = torch.nn.LSTM(p, p).double()
        hidden64 = (hidden[0].double(), hidden[1].double())

        def lstm_fn(x, h):                  # functional LSTM on the fp64 leaves so that they receive gradients
            return torch._VF.lstm(x, h, [P64["lstm.weight_ih_l0"], P64["lstm.weight_hh_l0"], P64["lstm.bias_ih_l0"],
                                         P64["lstm.bias_hh_l0"]], True, 1, 0.0, False, False, False)[0::1][0], None
        lstm64 = lambda x, h: (lstm_fn(x, h)[0], None)
    prob64, _ = R.ppo_pi_dense(P64, nfm.double(), ei, reachable, T, lstm64, hidden64)
    v64, _ = R.ppo_v_dense(P64, nfm.double(), ei, reachable, T, critic, lstm64, hidden64)
    ((prob64 * wprob.double()).sum() + (v64 * wv.double()).sum()).backward()
    g64 = {k: (P64[k].grad.detach() if P64[k].grad is not None else torch.zeros_like(P64[k])) for k in grads}
    own = max(float((grads[k].double() - g64[k]).abs().max()) / _scale_of(g64[k]) for k in grads if _scale_of(g64[k]) > 1e-6)
    print("  %-28s reference's own fp32 error vs fp64: prob %.1e v %.1e grads %.1e" % (
        "", float((prob.detach().double() - prob64.detach()).abs().max()),
        float((v.detach().double() - v64.detach()).abs().max()) / _scale_of(v64), own))
    print("  %-22s prob max %.3g  v scale %.3g" % (name, float(prob.max()), _scale_of(v)))
    meta = dict(kind="ppo", S=S, N=N, emb_dim=p, T=T, lstm_type=lstm_type, critic=critic, seed=seed, topo=topo.name)
    inp = dict(nfm=_np(nfm), edge_index=_np(ei).astype(np.int32), reachable=_np(reachable),
               wprob=_np(wprob), wv=_np(wv))
    if hidden is not None:
        inp["h0"], inp["c0"] = _np(hidden[0]), _np(hidden[1])
    _save(name, meta, P={k: _np(v_) for k, v_ in P.items()}, inp=inp,
          out=dict(prob=_np(prob), v=_np(v), mu=_np(mu), mu_s2v=_np(mu_s2v), greedy=_np(greedy),
                   prob_f64=_np(prob64), v_f64=_np(v64)),
          grad={k: _np(g) for k, g in grads.items()},
          grad64={k: _np(g) for k, g in g64.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    co, mb = load_reference()
    gen = torch.Generator().manual_seed(1234)

    print("DQN cases")
    m5 = graphs.load_topology("Manhattan5")
    x5 = graphs.synth_nfm(m5, 4, gen)
    dqn_case(co, "dqn_manhattan5_b4", [25] * 4, [m5.edge_index] * 4, 25, 64, 5, True, seed=0, xs=list(x5))
    dqn_case(co, "dqn_manhattan5_b4_sumpool_x01", [25] * 4, [m5.edge_index] * 4, 25, 64, 5, False, seed=1, scale=0.1,
             xs=list(x5))
    # variable-size directed graphs padded beyond their size (M3M5Mix-like; pins row/column conventions + Npad)
    sizes = [5, 9, 1, 12, 7]
    eis = [_random_digraph(s, 2.5, gen) for s in sizes]
    dqn_case(co, "dqn_varsize_directed_pad14", sizes, eis, 14, 16, 3, True, seed=2)
    dqn_case(co, "dqn_varsize_directed_T1", sizes, eis, 12, 24, 1, False, seed=3)
    m3 = graphs.load_topology("Manhattan3")
    xm = [graphs.synth_nfm(m3, 1, gen)[0], graphs.synth_nfm(m5, 1, gen)[0], graphs.synth_nfm(m3, 1, gen)[0]]
    dqn_case(co, "dqn_m3m5mix_pad25", [9, 25, 9], [m3.edge_index, m5.edge_index, m3.edge_index], 25, 64, 5, True,
             seed=4, xs=xm)
    ams = graphs.load_topology("NWB_AMS")
    xa = graphs.synth_nfm(ams, 2, gen)
    dqn_case(co, "dqn_ams_b2", [975] * 2, [ams.edge_index] * 2, 975, 64, 5, True, seed=5, xs=list(xa))
    dqn_case(co, "dqn_ams_b2_x01", [975] * 2, [ams.edge_index] * 2, 975, 64, 5, True, seed=6, scale=0.1, xs=list(xa))

    print("PPO cases")
    mem = graphs.load_topology("MemTask")
    ppo_case(mb, "ppo_memtask_p24_none_q", mem, 5, 24, 5, "None", "q", seed=10, obs_mask_freq=5)
    ppo_case(mb, "ppo_memtask_p64_emb_q", mem, 5, 64, 5, "EMB", "q", seed=11, obs_mask_freq=5)
    ppo_case(mb, "ppo_memtask_p24_none_v", mem, 5, 24, 5, "None", "v", seed=12)
    ppo_case(mb, "ppo_manhattan5_p64_emb_q", m5, 6, 64, 5, "EMB", "q", seed=13)
    ppo_case(mb, "ppo_ams_s3_p64_none_q", ams, 3, 64, 5, "None", "q", seed=14)


if __name__ == "__main__":
    main()
