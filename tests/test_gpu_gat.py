"""GPU parity tests of the GATv2 branch: the CUDA path (edge-attention kernels behind
s2v_gat_attn_forward/_backward + library GEMMs for lin_l/lin_r) against the goldens produced by the reference's own
QNet_GATv2 / PPO_GNN_Single_LSTM(qnet='gat2') classes and against the CPU oracle (oracle/gat_ref.py).

Tolerance: 1e-5 of the tensor scale; where the reference's own fp32 rounding error against the fp64 restatement is
larger (the attention gradients are differences of nearly equal terms: att / lin_r gradients have scales of 1e-5..1e-8
of the others) the bound is 2x the reference's own error -- the same rule as tests/test_gpu_parity.py.
"""
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_cases, load_golden
from test_gpu_parity import TOL, _check

pytestmark = pytest.mark.gpu


def _t(a, dev, dtype=None):
    t = torch.from_numpy(np.asarray(a))
    return t.to(dev) if dtype is None else t.to(dev, dtype)


@pytest.mark.parametrize("case", golden_cases("gat_dqn"))
def test_gat_dqn_golden(case, cuda_device):
    """QNet_GATv2.forward(xv, Ws, pyg_data): q, loss and every parameter gradient."""
    import simmodel_b200 as sb
    meta, G = load_golden(case)
    dev = cuda_device
    net = sb.QNet_GATv2({"emb_dim": meta["emb_dim"], "emb_iter_T": 5, "norm_agg": meta["norm_agg"], "node_dim": 7}, verbose=False)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in G["P"].items()}, strict=True)
    net = net.to(dev)
    ptr = G["inp"]["ptr"].astype(np.int64)
    ei = G["inp"]["edge_index"].astype(np.int64)
    datas = []
    for b in range(len(meta["sizes"])):
        m = (ei[0] >= ptr[b]) & (ei[0] < ptr[b + 1])
        datas.append(sb.Data(torch.from_numpy(G["inp"]["x"][ptr[b]:ptr[b + 1]]), torch.from_numpy(ei[:, m] - ptr[b])))
    batch = sb.Batch.from_data_list(datas)
    assert np.array_equal(batch.edge_index.numpy(), ei)
    q = net(None, None, batch)
    _check("q", q.detach().cpu().numpy(), G["out"]["q"], G["out"]["q_f64"])
    sel = _t(G["inp"]["actions"], dev) + _t(ptr, dev)[:-1]
    loss = torch.nn.SmoothL1Loss()(q[sel], _t(G["inp"]["targets"], dev))
    loss.backward()
    _check("loss", loss.detach().cpu().numpy(), G["out"]["loss"])
    for k, prm in net.named_parameters():
        _check("grad " + k, prm.grad.cpu().numpy(), G["grad"][k], G["grad64"][k])


def _ppo_model(meta, P, dev):
    import simmodel_b200 as sb
    config = {"lstm_type": meta["lstm_type"], "qnet": "gat2", "critic": meta["critic"], "emb_dim": meta["emb_dim"],
              "emb_iterT": meta["T"], "gat_concat": True, "gat_heads": 4, "gat_share_weights": False}
    hp = types.SimpleNamespace(node_dim=7, emb_dim=meta["emb_dim"], parallel_rollouts=4, batch_size=2, learning_rate=1e-3)
    model = sb.PPO_GNN_Single_LSTM(config, hp, {"base_checkpoint_path": None}, device=dev)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in P.items()}, strict=True)
    return model


@pytest.mark.parametrize("case", golden_cases("gat_ppo"))
def test_gat_ppo_golden(case, cuda_device):
    """pi / v / create_embeddings / greedy action of PPO_GNN_Single_LSTM(qnet='gat2') incl. the reference's shipped
    trained checkpoint on the real AMS graph."""
    meta, G = load_golden(case)
    dev = cuda_device
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    if "P" in G:
        P = G["P"]
    else:
        z = np.load(os.path.join(GOLDEN, "gat2_ppo_checkpoint.npz"))
        P = {k: z[k] for k in z.files}
    model = _ppo_model(meta, P, dev)
    nfm = torch.from_numpy(G["inp"]["nfm"])
    ei = torch.from_numpy(G["inp"]["edge_index"].astype(np.int64))
    reach = torch.from_numpy(G["inp"]["reachable"])
    hidden = None
    if "h0" in G["inp"]:
        hidden = (_t(G["inp"]["h0"], dev), _t(G["inp"]["c0"], dev))
    prob, _ = model.pi(nfm, ei, reach, hidden)
    v, _ = model.v(nfm, ei, reach, hidden)
    assert prob.shape == G["out"]["prob"].shape and v.shape == G["out"]["v"].shape
    tol = 1e-4 if meta["lstm_type"] == "EMB" else TOL       # cuDNN LSTM between our embedding and our head
    _check("prob", prob.detach().cpu().numpy(), G["out"]["prob"], G["out"]["prob_f64"], tol=tol)
    _check("v", v.detach().cpu().numpy(), G["out"]["v"], G["out"]["v_f64"], tol=tol)
    assert np.array_equal(prob.detach().cpu().numpy() == 0.0, G["out"]["prob"] == 0.0)
    loss = (prob * _t(G["inp"]["wprob"], dev)).sum() + (v * _t(G["inp"]["wv"], dev)).sum()
    loss.backward()
    for k, prm in model.named_parameters():
        g = prm.grad.cpu().numpy() if prm.grad is not None else np.zeros(tuple(prm.shape), dtype=np.float32)
        _check("grad " + k, g, G["grad"][k], G["grad64"][k], tol=tol, abs_floor=2e-7)
    # the embedding alone is ours end to end: 1e-5 in every case
    model.config = dict(model.config, lstm_type="None")
    mu, _ = model.create_embeddings(meta["S"], nfm, ei, None)
    _check("mu_gat", mu.detach().cpu().numpy(), G["out"]["mu_gat"])
    model.config = dict(model.config, lstm_type=meta["lstm_type"])
    arg, _ = model.greedy_action(nfm, ei, reach, hidden)
    assert np.array_equal(arg.cpu().numpy(), G["out"]["greedy"])


@pytest.mark.parametrize("heads,C,n,deg,seed", [(4, 32, 300, 3.0, 0), (2, 32, 257, 12.0, 1), (4, 16, 64, 2.0, 2),
                                                (4, 12, 100, 9.0, 3), (4, 6, 50, 3.0, 4), (1, 30, 40, 5.0, 5),
                                                (2, 8, 33, 20.0, 6), (1, 128, 20, 2.0, 7)])
def test_gat_conv_vs_oracle(heads, C, n, deg, seed, cuda_device):
    """One GATv2Conv layer (forward, input gradient, all parameter gradients) on random directed graphs with self
    loops, isolated nodes and in-degrees on both sides of the register-resident limit, every lane mapping."""
    import simmodel_b200 as sb
    from oracle import gat_ref as R
    dev = cuda_device
    gen = torch.Generator().manual_seed(seed)
    cin = 24
    ei = (torch.rand(n, n, generator=gen) < deg / n).nonzero().t().contiguous()
    x = torch.randn(n, cin, generator=gen)
    conv = sb.GATv2Conv(cin, C, heads=heads)
    with torch.no_grad():
        conv.bias.copy_(torch.randn(heads * C, generator=gen) * 0.1)
    P = {k: v.detach().clone().double().requires_grad_(True) for k, v in conv.state_dict().items()}
    x64 = x.double().requires_grad_(True)
    w = torch.randn(n, heads * C, generator=gen)
    for relu in (False, True):
        ref = R.gatv2_conv(x64, ei, P["lin_l.weight"], P["lin_l.bias"], P["lin_r.weight"], P["lin_r.bias"], P["att"], P["bias"])
        ref = torch.relu(ref) if relu else ref
        for t in list(P.values()) + [x64]:
            t.grad = None
        (ref * w.double()).sum().backward()
        conv = conv.to(dev)
        conv.zero_grad()
        xd = x.to(dev).requires_grad_(True)
        layout = sb.gat.gat_layout_from_batch(ei.to(dev), torch.tensor([0, n], device=dev))
        out = conv.forward_layout(xd, layout, relu=relu)
        (out * w.to(dev)).sum().backward()
        _check("out", out.detach().cpu().numpy(), ref.detach().numpy(), tol=3e-6)
        _check("dx", xd.grad.cpu().numpy(), x64.grad.numpy(), tol=3e-6)
        for k, prm in conv.named_parameters():
            _check("grad " + k, prm.grad.cpu().numpy(), P[k].grad.numpy(), tol=3e-6)


@pytest.mark.parametrize("heads,C", [(2, 32), (4, 12)])
def test_gat_conv_shared_weights_vs_oracle(heads, C, cuda_device):
    """share_weights=True (models_ngo.py:241-249: x_r = x_l = lin_l(x)): forward, input and parameter gradients."""
    import simmodel_b200 as sb
    from oracle import gat_ref as R
    dev = cuda_device
    gen = torch.Generator().manual_seed(11)
    n, cin = 120, 24
    ei = (torch.rand(n, n, generator=gen) < 4.0 / n).nonzero().t().contiguous()
    x = torch.randn(n, cin, generator=gen)
    conv = sb.GATv2Conv(cin, C, heads=heads, share_weights=True)
    assert conv.lin_r is conv.lin_l and "lin_r.weight" in conv.state_dict()
    P = {k: v.detach().clone().double().requires_grad_(True) for k, v in conv.state_dict().items() if not k.startswith("lin_r")}
    x64 = x.double().requires_grad_(True)
    w = torch.randn(n, heads * C, generator=gen)
    ref = R.gatv2_conv(x64, ei, P["lin_l.weight"], P["lin_l.bias"], P["lin_l.weight"], P["lin_l.bias"], P["att"], P["bias"])
    (torch.relu(ref) * w.double()).sum().backward()
    conv = conv.to(dev)
    xd = x.to(dev).requires_grad_(True)
    layout = sb.gat.gat_layout_from_batch(ei.to(dev), torch.tensor([0, n], device=dev))
    out = conv.forward_layout(xd, layout, relu=True)
    (out * w.to(dev)).sum().backward()
    _check("out", out.detach().cpu().numpy(), torch.relu(ref).detach().numpy(), tol=3e-6)
    _check("dx", xd.grad.cpu().numpy(), x64.grad.numpy(), tol=3e-6)
    for k, prm in conv.named_parameters():
        _check("grad " + k, prm.grad.cpu().numpy(), P[k].grad.numpy(), tol=3e-6)


def test_gat_replicated_equals_batched_and_bit_stable(cuda_device):
    """S copies of one topology: replicated layout (PPO path) == batched layout (DQN path) bit for bit, run-to-run
    bit-stable gradients (no atomics)."""
    import simmodel_b200 as sb
    dev = cuda_device
    topo = sb.graphs.load_topology("NWB_AMS")
    S = 6
    gen = torch.Generator().manual_seed(0)
    x = sb.graphs.synth_nfm(topo, S, gen).reshape(-1, 7).to(dev)
    torch.manual_seed(0)
    gat = sb.GATv2(in_channels=7, hidden_channels=128, num_layers=5, out_channels=64, heads=4).to(dev)
    rep = sb.gat.gat_layout_replicated(topo.edge_index.to(dev), topo.num_nodes, S)
    ptr = torch.arange(S + 1, device=dev) * topo.num_nodes
    bat = sb.gat.gat_layout_from_batch(sb.graphs.batch_edge_index(topo, S).to(dev), ptr)
    w = torch.randn(S * topo.num_nodes, 64, generator=torch.Generator().manual_seed(1)).to(dev)
    grads = []
    for layout in (rep, bat, rep):
        gat.zero_grad()
        mu = gat.forward_layout(x, layout)
        (mu * w).sum().backward()
        grads.append((mu.detach().clone(), [p.grad.clone() for p in gat.parameters()]))
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][0], grads[2][0])
    for a, b, c, (k, _) in zip(grads[0][1], grads[1][1], grads[2][1], gat.named_parameters()):
        if "att" in k or k.endswith("convs.0.bias") or ".bias" in k and "lin_" not in k:
            assert torch.equal(a, c), k                    # our fixed-order reductions: bit-stable run to run
    # a graph alone == the same graph inside the batch (rows are independent of the batch around them)
    one = sb.gat.gat_layout_replicated(topo.edge_index.to(dev), topo.num_nodes, 1)
    mu1 = gat.forward_layout(x[: topo.num_nodes].contiguous(), one)
    # (975 rows run the fp32 library GEMMs, 5 850 rows the tcgen05 bf16x3 lin kernels: two arithmetic paths, both within
    # the parity bound of each other)
    _check("graph alone", mu1.detach().cpu().numpy(), grads[0][0][: topo.num_nodes].cpu().numpy(), tol=1e-5)


def test_gat_no_cpu_path():
    import simmodel_b200 as sb
    conv = sb.GATv2Conv(7, 16, heads=2)
    with pytest.raises(sb.S2VError):
        conv(torch.zeros(4, 7), torch.zeros(2, 0, dtype=torch.int64))


def test_qfunction_with_gat_qnet(cuda_device):
    """QFunction.predict / get_best_action / batch_update drive QNet_GATv2 the way comb_opt.py does for --qnet gat2:
    pyg_data = Data(x=nfm, edge_index=EI) (comb_opt.py:472,522), state_tsr / W are dummies."""
    import simmodel_b200 as sb
    meta, G = load_golden("gat_dqn_manhattan5_b4")
    dev = cuda_device
    net = sb.QNet_GATv2({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in G["P"].items()}, strict=True)
    net = net.to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    qf = sb.QFunction(net, opt, torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.99999))
    ptr = G["inp"]["ptr"].astype(np.int64)
    ei = G["inp"]["edge_index"].astype(np.int64)
    zerovec = torch.zeros(1)
    datas = []
    for b in range(4):
        m = (ei[0] >= ptr[b]) & (ei[0] < ptr[b + 1])
        datas.append(sb.Data(torch.from_numpy(G["inp"]["x"][ptr[b]:ptr[b + 1]]), torch.from_numpy(ei[:, m] - ptr[b])))
    q0 = qf.predict(zerovec, zerovec, datas[0])
    # a graph alone == the same graph inside the golden batch, up to the library GEMM's batch-dependent rounding
    _check("predict", q0.cpu().numpy(), G["out"]["q"][:25], tol=2e-5)
    reach = [1, 5, 6]
    a, node, val = qf.get_best_action(zerovec, zerovec, datas[0], reach)
    ref = q0.cpu().numpy()[reach]
    assert a == int(np.argmax(ref)) and node == reach[a] and val == ref[a]
    loss = qf.batch_update([zerovec] * 4, [zerovec] * 4, datas, G["inp"]["actions"].tolist(), G["inp"]["targets"].tolist())
    assert abs(loss - float(G["out"]["loss"])) <= 1e-5 * max(1.0, abs(float(G["out"]["loss"])))


@pytest.mark.parametrize("n", [64, 1000, 5000])
def test_tensor_core_lin_kernels(n, cuda_device):
    """s2v_rows_gemm64 / s2v_rows_reduce64 (tcgen05, bf16x3) against fp64: forward x [Wl;Wr]^T, backward dx with and
    without the ReLU mask, weight gradient dxlr^T x; ragged last tile; run-to-run bit stability."""
    import ctypes as C
    from simmodel_b200 import _lib
    from simmodel_b200.layout import _p, _stream
    L = _lib.lib()
    dev = cuda_device
    gen = torch.Generator().manual_seed(n)
    x = torch.randn(n, 64, generator=gen).to(dev)
    w = (torch.randn(128, 64, generator=gen) * 0.2).to(dev)
    dy = torch.randn(n, 128, generator=gen).to(dev)
    y = torch.empty(n, 128, device=dev)
    _lib.check(L.s2v_rows_gemm64(_p(x), 64, 1, _p(w), 2, 0, None, 0, _p(y), 128, n, _stream()), "fwd")
    _check("y", y.cpu().numpy(), (x.double() @ w.double().t()).cpu().numpy(), tol=2e-6)
    dx = torch.empty(n, 64, device=dev)
    _lib.check(L.s2v_rows_gemm64(_p(dy), 128, 2, _p(w), 1, 1, None, 0, _p(dx), 64, n, _stream()), "dx")
    ref = dy.double() @ w.double()
    _check("dx", dx.cpu().numpy(), ref.cpu().numpy(), tol=2e-6)
    _lib.check(L.s2v_rows_gemm64(_p(dy), 128, 2, _p(w), 1, 1, _p(x), 64, _p(dx), 64, n, _stream()), "dx masked")
    _check("dx masked", dx.cpu().numpy(), (ref * (x > 0)).cpu().numpy(), tol=2e-6)
    assert bool(((dx == 0) == ((x <= 0) | (ref.float() == 0) | (dx == 0))).all())
    ws = torch.empty(L.s2v_rows_reduce64_workspace_bytes(2), dtype=torch.uint8, device=dev)
    outs = []
    for _ in range(2):
        dw = torch.empty(128, 64, device=dev)
        _lib.check(L.s2v_rows_reduce64(_p(dy), 128, 2, _p(x), 64, _p(dw), _p(ws), ws.numel(), n, _stream()), "dw")
        outs.append(dw)
    assert torch.equal(outs[0], outs[1])
    _check("dw", outs[0].cpu().numpy(), (dy.double().t() @ x.double()).cpu().numpy(), tol=3e-6)
    assert L.s2v_rows_gemm64(_p(x), 64, 1, _p(w), 5, 0, None, 0, _p(y), 320, n, _stream()) == -2      # unsupported shape (> 256 outputs)


def test_gat_large_batch_tensor_lin_vs_oracle(cuda_device, monkeypatch):
    """>= 4096 rows: the hidden layers' lin_l / lin_r GEMMs run on the tcgen05 kernels (ReLU mask fused into the dx
    epilogue).  QNet_GATv2 on 4 UTR graphs against the CPU oracle (fp32, adjudicated in fp64); the all-cuBLAS path
    (S2V_B200_PATH=ffma) on the same inputs for comparison."""
    import simmodel_b200 as sb
    from oracle import gat_ref as R
    dev = cuda_device
    topo = sb.graphs.load_topology("NWB_UTR")
    B = 4
    gen = torch.Generator().manual_seed(5)
    x = sb.graphs.synth_nfm(topo, B, gen).reshape(-1, 7)
    ei = sb.graphs.batch_edge_index(topo, B)
    ptr = torch.arange(B + 1, dtype=torch.int64) * topo.num_nodes
    torch.manual_seed(1)
    net = sb.QNet_GATv2({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False)
    with torch.no_grad():
        for k, prm in net.named_parameters():
            if k.endswith(".bias") and "convs" in k and "lin_" not in k:
                prm.copy_(torch.randn(prm.shape) * 0.1)
    P = {k: v.detach().clone() for k, v in net.state_dict().items()}
    sel = torch.randint(0, topo.num_nodes, (B,), generator=gen) + ptr[:-1]
    tg = torch.randn(B, generator=gen)
    refs = {}
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        Pg = {k: v.detach().to(dt).clone().requires_grad_(True) for k, v in P.items()}
        q = R.qnet_gat_forward(Pg, x.to(dt), ei, ptr, 5, True)
        torch.nn.SmoothL1Loss()(q[sel], tg.to(dt)).backward()
        refs[name] = (q.detach(), {k: v.grad for k, v in Pg.items()})
    net = net.to(dev)
    layout = sb.gat.gat_layout_from_batch(ei.to(dev), ptr.to(dev))
    assert all(c.uses_tensor_lin(layout.num_nodes) for c in net.gat.convs)      # r02: the F = 7 input layer too
    for path, tol in (("auto", 1e-4), ("ffma", 2e-5)):
        monkeypatch.setenv("S2V_B200_PATH", path)
        net.zero_grad()
        q = net.forward_graph(x.to(dev), layout)
        torch.nn.SmoothL1Loss()(q[sel.to(dev)], tg.to(dev)).backward()
        _check("q " + path, q.detach().cpu().numpy(), refs["f32"][0].numpy(), refs["f64"][0].numpy(), tol=1e-5)
        for k, prm in net.named_parameters():
            _check("grad %s %s" % (k, path), prm.grad.cpu().numpy(), refs["f32"][1][k].numpy(), refs["f64"][1][k].numpy(), tol=tol)


@pytest.mark.parametrize("heads,C", [(2, 32), (4, 32), (4, 16), (4, 12)])
def test_gat_conv_edge_cases(heads, C, cuda_device):
    """No edges at all (every node attends to itself only: out = x_l + bias), and a hub with 40 incoming / outgoing
    edges (beyond the register-resident slots and beyond the ids a lane group prefetches), specialised and generic
    kernels, batched and replicated layouts."""
    import simmodel_b200 as sb
    from oracle import gat_ref as R
    dev = cuda_device
    gen = torch.Generator().manual_seed(heads * 100 + C)
    n, cin = 41, 16
    conv = sb.GATv2Conv(cin, C, heads=heads)
    with torch.no_grad():
        conv.bias.copy_(torch.randn(heads * C, generator=gen) * 0.1)
    P = {k: v.detach().clone().double().requires_grad_(True) for k, v in conv.state_dict().items()}
    conv = conv.to(dev)
    hub = torch.arange(1, n)
    eis = {"no_edges": torch.zeros(2, 0, dtype=torch.int64),
           "hub": torch.cat((torch.stack((hub, torch.zeros_like(hub))), torch.stack((torch.zeros_like(hub), hub)),
                             torch.tensor([[5, 7, 7], [6, 8, 7]])), dim=1)}
    for name, ei in eis.items():
        for copies in (1, 3):
            x = torch.randn(copies * n, cin, generator=gen)
            w = torch.randn(copies * n, heads * C, generator=gen)
            x64 = x.double().requires_grad_(True)
            for t in P.values():
                t.grad = None
            ei_b = torch.cat([ei + k * n for k in range(copies)], dim=1)
            ref = R.gatv2_conv(x64, ei_b, P["lin_l.weight"], P["lin_l.bias"], P["lin_r.weight"], P["lin_r.bias"], P["att"], P["bias"])
            (ref * w.double()).sum().backward()
            layouts = [sb.gat.gat_layout_from_batch(ei_b.to(dev), torch.arange(copies + 1, device=dev) * n)]
            if copies > 1:
                layouts.append(sb.gat.gat_layout_replicated(ei.to(dev), n, copies))
            for layout in layouts:
                conv.zero_grad()
                xd = x.to(dev).requires_grad_(True)
                out = conv.forward_layout(xd, layout)
                (out * w.to(dev)).sum().backward()
                _check("%s out" % name, out.detach().cpu().numpy(), ref.detach().numpy(), tol=3e-6)
                _check("%s dx" % name, xd.grad.cpu().numpy(), x64.grad.numpy(), tol=3e-6)
                for k, prm in conv.named_parameters():
                    _check("%s grad %s" % (name, k), prm.grad.cpu().numpy(), P[k].grad.numpy(), tol=3e-6, abs_floor=1e-12)


@pytest.mark.parametrize("heads,C", [(2, 32), (4, 32), (4, 12)])
def test_gat_kernels_stay_inside_their_buffers(heads, C, cuda_device):
    """Every output / workspace buffer of the GAT entry points is surrounded by canary zones that must survive the
    forward and the backward (specialised per-edge-alpha kernels and the generic statistics kernels; ragged row counts,
    a hub node beyond the register-resident slots)."""
    import ctypes as Ct
    import simmodel_b200 as sb
    from simmodel_b200 import _lib
    from simmodel_b200.layout import _p, _stream
    L = _lib.lib()
    dev = cuda_device
    gen = torch.Generator().manual_seed(3)
    n, HC, GUARD = 1237, heads * C, 4096
    hub = torch.arange(1, 41)
    ei = torch.cat(((torch.rand(n, n, generator=gen) < 3.0 / n).nonzero().t(), torch.stack((hub, torch.zeros_like(hub))),
                    torch.stack((torch.zeros_like(hub), hub))), dim=1)
    layout = sb.gat.gat_layout_from_batch(ei.to(dev), torch.tensor([0, n], device=dev))
    g = layout.c_struct()
    edge_alpha = bool(L.s2v_gat_uses_edge_alpha(heads, C, 2 * HC))
    ef = int(L.s2v_gat_edge_floats(Ct.byref(g), heads)) if edge_alpha else 0

    class Guarded:
        def __init__(self, numel, dtype=torch.float32):
            self.full = torch.full((numel + 2 * GUARD,), 12345.0 if dtype == torch.float32 else 77, dtype=dtype, device=dev)
            self.view = self.full[GUARD:GUARD + numel]
            self.fill = self.full[0].item()

        def intact(self):
            return bool((self.full[:GUARD] == self.fill).all()) and bool((self.full[-GUARD:] == self.fill).all())
    xlr = torch.randn(n, 2 * HC, generator=gen).to(dev)
    att, bias, lb = (torch.randn(k, generator=gen).to(dev) * 0.1 for k in (HC, HC, 2 * HC))
    dout = torch.randn(n, HC, generator=gen).to(dev)
    out, keep = Guarded(n * HC), Guarded(ef if edge_alpha else n * heads * 4)
    dxlr, gab = Guarded(n * 2 * HC), Guarded(4 * HC)
    ws = Guarded(int(L.s2v_gat_attn_backward_workspace_bytes(heads, C, ef)), torch.uint8)
    xr_p = Ct.c_void_p(xlr.data_ptr() + 4 * HC)
    _lib.check(L.s2v_gat_attn_forward(Ct.byref(g), heads, C, 2 * HC, _p(xlr), xr_p, _p(att), _p(bias), _p(lb), 1, _p(out.view),
                                      None if edge_alpha else _p(keep.view), _p(keep.view) if edge_alpha else None, _stream()), "fwd")
    _lib.check(L.s2v_gat_attn_backward(Ct.byref(g), heads, C, 2 * HC, _p(xlr), xr_p, _p(att), _p(lb),
                                       None if edge_alpha else _p(keep.view), _p(keep.view) if edge_alpha else None,
                                       _p(layout.row_to_col), _p(dout), _p(dxlr.view), Ct.c_void_p(dxlr.view.data_ptr() + 4 * HC),
                                       _p(gab.view[:HC]), _p(gab.view[HC:2 * HC]), _p(gab.view[2 * HC:]), _p(ws.view), ws.view.numel(),
                                       _stream()), "bwd")
    torch.cuda.synchronize()
    for name, b in (("out", out), ("alpha/stats", keep), ("dxlr", dxlr), ("grads", gab), ("workspace", ws)):
        assert b.intact(), name + " canary overwritten"
    assert bool(torch.isfinite(out.view).all()) and bool(torch.isfinite(dxlr.view).all()) and bool(torch.isfinite(gab.view).all())
    if edge_alpha:                                                     # every alpha slot was written: rows sum to one per head
        assert bool((keep.view >= 0).all()) and bool((keep.view <= 1.0 + 1e-6).all())


def test_lin_and_head_backward_kernels_stay_inside_their_buffers(cuda_device):
    """Canary zones around the outputs of the tensor-core lin kernels (ragged last tile) and of the head backward's
    streaming + per-row passes (sparse dq, >= 4096 rows)."""
    import simmodel_b200 as sb
    from simmodel_b200 import _lib, ops
    from simmodel_b200.layout import _p, _stream
    L = _lib.lib()
    dev = cuda_device
    GUARD = 4096
    gen = torch.Generator().manual_seed(9)

    def guarded(numel):
        full = torch.full((numel + 2 * GUARD,), 12345.0, device=dev)
        return full, full[GUARD:GUARD + numel]

    def intact(full):
        return bool((full[:GUARD] == 12345.0).all()) and bool((full[-GUARD:] == 12345.0).all())
    n = 4999
    x, w, dy = torch.randn(n, 64, generator=gen).to(dev), (torch.randn(128, 64, generator=gen) * 0.2).to(dev), torch.randn(n, 128, generator=gen).to(dev)
    fy, y = guarded(n * 128)
    fdx, dx = guarded(n * 64)
    fdw, dw = guarded(128 * 64)
    fws, ws = guarded(int(L.s2v_rows_reduce64_workspace_bytes(2)) // 4)
    _lib.check(L.s2v_rows_gemm64(_p(x), 64, 1, _p(w), 2, 0, None, 0, _p(y), 128, n, _stream()), "fwd")
    _lib.check(L.s2v_rows_gemm64(_p(dy), 128, 2, _p(w), 1, 1, _p(x), 64, _p(dx), 64, n, _stream()), "dx")
    _lib.check(L.s2v_rows_reduce64(_p(dy), 128, 2, _p(x), 64, _p(dw), _p(ws), ws.numel() * 4, n, _stream()), "dw")
    torch.cuda.synchronize()
    assert intact(fy) and intact(fdx) and intact(fdw) and intact(fws)
    # head backward, sparse dq on 5 UTR graphs (5910 rows): stream + per-row passes write dmu
    topo = sb.graphs.load_topology("NWB_UTR")
    B = 5
    layout = sb.GraphLayout.replicated(topo.edge_index.to(dev), topo.num_nodes, B)
    nn_ = layout.num_nodes
    torch.manual_seed(0)
    net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False).to(dev)
    mu = torch.randn(nn_, 64, generator=gen).to(dev)
    hw = [t.detach() for t in net.head_weights()]
    pool, gstate, q = ops._head_forward(layout, True, mu, hw)
    dq = torch.zeros(nn_, device=dev)
    dq[torch.tensor([0, 63, 64, 1181, 1182, nn_ - 1], device=dev)] = torch.randn(6, generator=gen).to(dev)
    fdmu, dmu = guarded(nn_ * 64)
    ops._head_backward(layout, True, mu, hw, pool, gstate, dq, True, dmu_out=dmu.view(nn_, 64))
    torch.cuda.synchronize()
    assert intact(fdmu) and bool(torch.isfinite(dmu).all())


@pytest.mark.parametrize("k,m", [(7, 128), (7, 256), (64, 128), (128, 256), (128, 128), (128, 64), (100, 192)])
def test_rows_linear_outer_general_shapes(k, m, cuda_device):
    """s2v_rows_linear / s2v_rows_outer (tcgen05, bf16x3) for every lin_l / lin_r shape of the shipped configurations --
    the F = 7 input layers (ragged K, zero padded), the 128-wide hidden layers of the PPO gat2 model (two output passes,
    two accumulating K passes in the backward) -- against fp64: forward, dx with and without the ReLU mask, dW."""
    import ctypes as C
    from simmodel_b200 import _lib
    from simmodel_b200.layout import _p, _stream
    L = _lib.lib()
    dev = cuda_device
    n = 5000 + k                      # not a multiple of the tile height
    gen = torch.Generator().manual_seed(k * 1000 + m)
    x = torch.randn(n, k, generator=gen).to(dev)
    w = (torch.randn(m, k, generator=gen) / 8).to(dev)
    dy = torch.randn(n, m, generator=gen).to(dev)

    def err(a, b):
        return float((a.double() - b).abs().max() / b.abs().max())
    y = torch.empty(n, m, device=dev)
    _lib.check(L.s2v_rows_linear(_p(x), k, k, _p(w), k, m, 0, None, 0, _p(y), m, n, _stream()), "fwd")
    assert err(y, x.double() @ w.double().t()) < 2e-6
    ws = torch.empty(L.s2v_rows_reduce64_workspace_bytes(2), dtype=torch.uint8, device=dev)
    dw = torch.empty(m, k, device=dev)
    _lib.check(L.s2v_rows_outer(_p(dy), m, m, _p(x), k, k, _p(dw), k, _p(ws), ws.numel(), n, _stream()), "dW")
    assert err(dw, dy.double().t() @ x.double()) < 2e-6
    if k % 64 == 0:
        ref = dy.double() @ w.double()
        dx = torch.empty(n, k, device=dev)
        _lib.check(L.s2v_rows_linear(_p(dy), m, m, _p(w), k, k, 1, None, 0, _p(dx), k, n, _stream()), "dx")
        assert err(dx, ref) < 2e-6
        _lib.check(L.s2v_rows_linear(_p(dy), m, m, _p(w), k, k, 1, _p(x), k, _p(dx), k, n, _stream()), "dx masked")
        assert err(dx, ref * (x > 0).double()) < 2e-6
        assert bool((dx[x <= 0] == 0).all())
