"""GATv2 branch, CPU side: the functional oracle (oracle/gat_ref.py) against the golden
vectors produced by the reference's own QNet_GATv2 / PPO_GNN_Single_LSTM(qnet='gat2') classes
(oracle/gen_golden_gat.py), and the host mirror's parameter layout against the reference's shipped checkpoint."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_cases, load_golden
from oracle import gat_ref as G


def _scale(a):
    return float(np.abs(a).max()) + 1e-30


@pytest.mark.parametrize("case", golden_cases("gat_dqn"))
def test_gat_oracle_reproduces_reference_dqn(case):
    meta, Z = load_golden(case)
    P = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in Z["P"].items()}
    x = torch.from_numpy(Z["inp"]["x"])
    ei = torch.from_numpy(Z["inp"]["edge_index"].astype(np.int64))
    ptr = torch.from_numpy(Z["inp"]["ptr"].astype(np.int64))
    q = G.qnet_gat_forward(P, x, ei, ptr, 5, meta["norm_agg"]).reshape(-1)
    assert np.abs(q.detach().numpy() - Z["out"]["q"]).max() <= 2e-6 * _scale(Z["out"]["q"])
    sel = torch.from_numpy(Z["inp"]["actions"]) + ptr[:-1]
    torch.nn.SmoothL1Loss()(q[sel], torch.from_numpy(Z["inp"]["targets"])).backward()
    for k in Z["grad"]:
        g64 = Z["grad64"][k]
        own = np.abs(Z["grad"][k] - g64).max() / _scale(g64)
        assert np.abs(P[k].grad.numpy() - g64).max() / _scale(g64) <= max(3e-6, 2 * own), k


@pytest.mark.parametrize("case", [c for c in golden_cases("gat_ppo") if "checkpoint" not in c])
def test_gat_oracle_reproduces_reference_ppo_embedding(case):
    meta, Z = load_golden(case)
    P = {k: torch.from_numpy(v) for k, v in Z["P"].items()}
    S, N, p = meta["S"], meta["N"], meta["emb_dim"]
    nfm = torch.from_numpy(Z["inp"]["nfm"])
    ei = torch.from_numpy(Z["inp"]["edge_index"].astype(np.int64))
    bt = G.batch_from_data_list([G.Data(e, ei) for e in nfm])
    mu = G.gat_forward(P, bt.x, bt.edge_index, meta["T"]).reshape(S, N, p).numpy()
    assert np.abs(mu - Z["out"]["mu_gat"]).max() <= 2e-6 * _scale(Z["out"]["mu_gat"])
    if meta["lstm_type"] == "None":
        prob, v = G.ppo_heads(P, torch.from_numpy(mu), torch.from_numpy(Z["inp"]["reachable"]), meta["critic"])
        assert np.abs(prob.numpy() - Z["out"]["prob"]).max() <= 2e-6
        assert np.abs(v.numpy() - Z["out"]["v"]).max() <= 2e-6 * _scale(Z["out"]["v"])


def test_self_loop_handling_and_edge_order():
    """remove_self_loops + add_self_loops: existing loops are dropped, one loop per node is appended last."""
    ei = torch.tensor([[0, 1, 1, 2, 2], [1, 1, 0, 2, 0]])
    out = G.remove_then_add_self_loops(ei, 4)
    assert out.tolist() == [[0, 1, 2, 0, 1, 2, 3], [1, 0, 0, 0, 1, 2, 3]]


def test_layer_dims_match_shipped_checkpoint():
    """Parameter names and shapes of the PPO gat2 model == the reference's shipped trained checkpoint."""
    z = np.load(os.path.join(GOLDEN, "gat2_ppo_checkpoint.npz"))
    dims = G.gat_layer_dims(7, 128, 5, 64, 4)
    assert dims == [(7, 128, 4), (128, 128, 4), (128, 128, 4), (128, 128, 4), (128, 64, 4)]
    for k, (cin, w, h) in enumerate(dims):
        assert z["gat.convs.%d.att" % k].shape == (1, h, w // h)
        assert z["gat.convs.%d.bias" % k].shape == (w,)
        assert z["gat.convs.%d.lin_l.weight" % k].shape == (w, cin) and z["gat.convs.%d.lin_r.weight" % k].shape == (w, cin)


def test_host_mirror_loads_shipped_checkpoint_strict():
    """simmodel_b200.PPO_GNN_Single_LSTM(qnet='gat2') has exactly the reference's state-dict keys, in its order
    (constructing and loading needs no GPU; running does)."""
    import simmodel_b200 as sb
    z = np.load(os.path.join(GOLDEN, "gat2_ppo_checkpoint.npz"))
    config = {"lstm_type": "EMB", "qnet": "gat2", "critic": "q", "emb_dim": 64, "emb_iterT": 5, "gat_concat": True,
              "gat_heads": 4, "gat_share_weights": False}
    hp = types.SimpleNamespace(node_dim=7, emb_dim=64, parallel_rollouts=4, batch_size=2, learning_rate=1e-3)
    model = sb.PPO_GNN_Single_LSTM(config, hp, {"base_checkpoint_path": None}, device="cpu")
    assert list(model.state_dict().keys()) == list(z.files)
    model.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=True)
    total, _ = model.numTrainableParameters()
    assert total == sum(int(np.prod(z[k].shape)) for k in z.files)
    with pytest.raises(sb.S2VError):                       # no CPU path
        model.pi(torch.zeros(8, 7), torch.zeros(2, 3, dtype=torch.int64), torch.ones(8, dtype=torch.bool), None)


def test_qnet_gatv2_mirror_keys_and_heads_fallback():
    import simmodel_b200 as sb
    net = sb.QNet_GATv2({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False)
    _, Z = load_golden("gat_dqn_manhattan5_b4")
    assert list(net.state_dict().keys()) == list(Z["P"].keys())
    net.load_state_dict({k: torch.from_numpy(v) for k, v in Z["P"].items()}, strict=True)
    # heads falls back to 1 for a width it does not divide (comb_opt.py:56-57)
    g = sb.GATv2(in_channels=7, hidden_channels=30, num_layers=3, out_channels=15, heads=4)
    assert [(c.heads, c.out_channels) for c in g.convs] == [(1, 30), (1, 30), (1, 15)]
    with pytest.raises(NotImplementedError):
        sb.GATv2Conv(7, 64, heads=4)                       # heads*channels > 128 is outside the kernels


@pytest.mark.parametrize("case", golden_cases("gat_dqn") + [c for c in golden_cases("gat_ppo") if "checkpoint" not in c])
def test_two_independent_gatv2_restatements_agree(case):
    """oracle/gat_ref.py (PyG's edge-list / segment-softmax form) vs oracle/gat_dense_ref.py (the paper's formula on a
    dense N x N neighbourhood matrix, written without reference to the first): the same embedding to fp64 rounding on
    every golden input -- self-loop handling, attention direction, bias placement and multi-edge weighting included."""
    from oracle import gat_dense_ref as D
    meta, Z = load_golden(case)
    P = {k: torch.from_numpy(v).double() for k, v in Z["P"].items()}
    if "x" in Z["inp"]:
        x = torch.from_numpy(Z["inp"]["x"]).double()
        ei = torch.from_numpy(Z["inp"]["edge_index"].astype(np.int64))
        layers = 5
    else:
        nfm = torch.from_numpy(Z["inp"]["nfm"])
        bt = G.batch_from_data_list([G.Data(e, torch.from_numpy(Z["inp"]["edge_index"].astype(np.int64))) for e in nfm])
        x, ei, layers = bt.x.double(), bt.edge_index, meta["T"]
    a = G.gat_forward(P, x, ei, layers)
    b = D.gat_forward_dense(P, x, ei, layers)
    assert float((a - b).abs().max()) <= 1e-11 * float(a.abs().max())


def test_dense_restatement_counts_multi_edges_and_ignores_given_self_loops():
    from oracle import gat_dense_ref as D
    ei = torch.tensor([[0, 0, 1, 2, 2, 3], [1, 1, 1, 0, 2, 0]])           # (0->1) twice, loops on 1 and 2
    M = D.neighbourhood_matrix(ei, 4)
    assert M.tolist() == [[1, 0, 1, 1], [2, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]
    torch.manual_seed(0)
    P = {"gat.convs.0." + k: v.double() for k, v in dict(att=torch.randn(1, 2, 4), bias=torch.randn(8), **{
        "lin_l.weight": torch.randn(8, 3), "lin_l.bias": torch.randn(8), "lin_r.weight": torch.randn(8, 3),
        "lin_r.bias": torch.randn(8)}).items()}
    x = torch.randn(4, 3).double()
    assert float((G.gat_forward(P, x, ei, 1) - D.gat_forward_dense(P, x, ei, 1)).abs().max()) <= 1e-12
