"""The oracle against the golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py, run in the build container).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden
from oracle import s2v_ref as R


def _scale(a):
    return float(np.abs(a).max()) + 1e-30


def _dense_inputs(meta, G):
    sizes, n_pad = meta["sizes"], meta["n_pad"]
    ptr = G["inp"]["ptr"].astype(np.int64)
    ei = G["inp"]["edge_index"].astype(np.int64)
    xv = torch.zeros(len(sizes), n_pad, 7)
    Ws = torch.zeros(len(sizes), n_pad, n_pad)
    for b, n in enumerate(sizes):
        xv[b, :n] = torch.from_numpy(G["inp"]["x"][ptr[b]:ptr[b + 1]])
        m = (ei[0] >= ptr[b]) & (ei[0] < ptr[b + 1])
        Ws[b, ei[0][m] - ptr[b], ei[1][m] - ptr[b]] = 1.0
    return xv, Ws


@pytest.mark.parametrize("case", [c for c in golden_cases("dqn") if "ams" not in c])
def test_dense_oracle_reproduces_reference(case):
    meta, G = load_golden(case)
    P = {k: torch.from_numpy(v) for k, v in G["P"].items()}
    xv, Ws = _dense_inputs(meta, G)
    bt = R.batch_from_sizes(meta["sizes"])
    q = R.qnet_forward_dense(P, xv, Ws, bt.ptr, bt.batch, meta["T"], meta["norm_agg"]).reshape(-1).numpy()
    # same torch ops in the same order; allow for a different CPU's vector paths
    assert np.abs(q - G["out"]["q"]).max() <= 2e-6 * _scale(G["out"]["q"])


@pytest.mark.parametrize("case", golden_cases("dqn"))
def test_sparse_oracle_matches_reference(case):
    meta, G = load_golden(case)
    P = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in G["P"].items()}
    x = torch.from_numpy(G["inp"]["x"])
    ei = torch.from_numpy(G["inp"]["edge_index"].astype(np.int64))
    ptr = torch.from_numpy(G["inp"]["ptr"].astype(np.int64))
    q = R.qnet_forward_sparse(P, x, ei, ptr, meta["n_pad"], meta["T"], meta["norm_agg"])
    loss = R.dqn_loss(q, G["inp"]["actions"].tolist(), ptr, G["inp"]["targets"].tolist())
    loss.backward()
    q64 = G["out"]["q_f64"]
    own = np.abs(G["out"]["q"] - q64).max() / _scale(q64)
    assert np.abs(q.detach().numpy() - q64).max() / _scale(q64) <= max(2e-6, 2 * own)
    assert abs(float(loss) - float(G["out"]["loss"])) <= 1e-5 * max(1.0, abs(float(G["out"]["loss"])))
    for k in G["grad"]:
        g64 = G["grad64"][k]
        own = np.abs(G["grad"][k] - g64).max() / _scale(g64)
        assert np.abs(P[k].grad.numpy() - g64).max() / _scale(g64) <= max(3e-6, 2 * own), k
    a, node, val = R.greedy_action(G["out"]["q"][: meta["sizes"][0]], meta["greedy"]["reachable"])
    # the golden greedy action was taken on graph 0 ALONE and unpadded; only check consistency of the helper
    assert node == meta["greedy"]["reachable"][a]


@pytest.mark.parametrize("case", [c for c in golden_cases("ppo") if "ams" not in c])
def test_ppo_dense_oracle_reproduces_reference(case):
    meta, G = load_golden(case)
    P = {k: torch.from_numpy(v) for k, v in G["P"].items()}
    nfm = torch.from_numpy(G["inp"]["nfm"])
    ei = torch.from_numpy(G["inp"]["edge_index"].astype(np.int64))
    reach = torch.from_numpy(G["inp"]["reachable"])
    lstm, hidden = None, None
    if meta["lstm_type"] == "EMB":
        lstm = torch.nn.LSTM(meta["emb_dim"], meta["emb_dim"])
        lstm.load_state_dict({k[5:]: v for k, v in P.items() if k.startswith("lstm.")})
        hidden = (torch.from_numpy(G["inp"]["h0"]), torch.from_numpy(G["inp"]["c0"]))
    with torch.no_grad():
        prob, _ = R.ppo_pi_dense(P, nfm, ei, reach, meta["T"], lstm, hidden)
        v, _ = R.ppo_v_dense(P, nfm, ei, reach, meta["T"], meta["critic"], lstm, hidden)
    assert np.abs(prob.numpy() - G["out"]["prob"]).max() <= 2e-6
    assert np.abs(v.numpy() - G["out"]["v"]).max() <= 2e-6 * _scale(G["out"]["v"])
    assert np.array_equal(prob.argmax(dim=1).numpy(), G["out"]["greedy"])


def test_csr_known_answer():
    ei = np.array([[2, 0, 2, 1, 0, 2], [1, 2, 0, 1, 0, 1]])
    rowptr, col, rowptr_t, col_t, indeg = R.csr_from_edge_index(ei, 4)
    assert rowptr.tolist() == [0, 2, 3, 6, 6]
    assert col.tolist() == [2, 0, 1, 1, 0, 1]          # edge order kept inside a row (duplicates too)
    assert rowptr_t.tolist() == [0, 2, 5, 6, 6]
    assert col_t.tolist() == [2, 0, 2, 1, 2, 0]
    assert indeg.tolist() == [2, 3, 1, 0]


def test_masked_helpers_known_answer():
    q = np.array([1.0, 3.0, 3.0, -1.0, 5.0, 5.0, 0.0], dtype=np.float32)
    ptr = np.array([0, 4, 7, 7])
    mask = np.array([1, 1, 1, 0, 0, 1, 1], dtype=bool)
    arg, val = R.masked_argmax_per_graph(q, mask, ptr)
    assert arg.tolist() == [1, 1, -1] and val[:2].tolist() == [3.0, 5.0] and np.isneginf(val[2])
    assert R.greedy_action(q, [3, 4, 5]) == (1, 4, np.float32(5.0))
    pr = R.masked_softmax_per_graph(torch.from_numpy(q), torch.from_numpy(mask), torch.from_numpy(ptr))
    assert pr[3] == 0 and pr[4] == 0 and abs(float(pr[:4].sum()) - 1) < 1e-6


def test_third_party_restatements():
    src = torch.arange(12, dtype=torch.float32).view(6, 2)
    idx = torch.tensor([0, 0, 2, 2, 2, 3])
    assert R.scatter_add(src, idx, 5).tolist() == [[2, 4], [0, 0], [18, 21], [10, 11], [0, 0]]
    assert R.scatter_mean(src, idx, 5)[1].tolist() == [0, 0]          # empty segment: 0 / clamp(0,1)
    assert R.scatter_mean(src, idx, 5)[2].tolist() == [6, 7]
    W = torch.tensor([[0., 1, 0], [1, 0, 1], [0, 0, 1]])
    ei, _ = R.dense_to_sparse(W)
    assert ei.tolist() == [[0, 1, 1, 2], [1, 0, 2, 2]]                # row-major
    assert torch.equal(R.to_dense_adj(ei)[0], W)
    b = R.batch_from_sizes([2, 3], [torch.tensor([[0], [1]]), torch.tensor([[0, 2], [1, 0]])])
    assert b.ptr.tolist() == [0, 2, 5] and b.batch.tolist() == [0, 0, 1, 1, 1]
    assert b.edge_index.tolist() == [[0, 2, 4], [1, 3, 2]]
