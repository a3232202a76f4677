"""GPU parity of the models_ngo mirrors (FeatureExtractor + Actor + Critic =
MaskablePPOPolicy) against goldens produced by the UNMODIFIED modules/ppo/models_ngo.py classes
(oracle/gen_golden_ngo.py): de-serialised features, action distribution over padded logits, values, all gradients,
the per-node LSTM state slots."""
import types

import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden
from test_gpu_parity import _check

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", golden_cases("ngo"))
def test_ngo_golden(case, cuda_device):
    import simmodel_b200 as sb
    meta, G = load_golden(case)
    dev = cuda_device
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    hp = types.SimpleNamespace(emb_dim=64, node_dim=7, lstm_on=meta["lstm_on"], hidden_size=64, recurrent_layers=1,
                               max_possible_nodes=meta["max_nodes"], critic=meta["critic"])
    pol = sb.ngo.MaskablePPOPolicy(None, meta["max_nodes"], False, False, init_log_std_dev=0.0, hp=hp)
    assert list(pol.state_dict().keys()) == list(G["P"].keys())
    pol.load_state_dict({k: torch.from_numpy(v) for k, v in G["P"].items()}, strict=True)
    pol = pol.to(dev)
    state = torch.from_numpy(G["inp"]["state"]).to(dev)
    feats, nodes_in_batch, _, num_nodes = pol.FE(state)
    assert nodes_in_batch == meta["nodes_in_batch"] and feats.shape == G["out"]["features"].shape
    f = feats.detach().cpu().numpy()
    assert np.array_equal(f[:, :, 64:], G["out"]["features"][:, :, 64:])            # batch ids, reachable flags, node counts
    _check("mu", f[:, :, :64], G["out"]["features"][:, :, :64], tol=1e-5)
    assert np.array_equal(num_nodes.cpu().numpy(), G["out"]["num_nodes"])
    dist, v = pol(feats)
    tol = 1e-4 if meta["lstm_on"] else 1e-5                                           # cuDNN LSTM between GAT and heads
    assert dist.probs.shape == G["out"]["probs"].shape and v.shape == G["out"]["v"].shape
    _check("probs", dist.probs.detach().numpy(), G["out"]["probs"], tol=tol)
    assert np.array_equal(dist.probs.detach().numpy() == 0.0, G["out"]["probs"] == 0.0)   # masked / padded entries
    _check("v", v.detach().cpu().numpy(), G["out"]["v"], tol=tol)
    loss = (dist.probs * torch.from_numpy(G["inp"]["wp"])).sum() + (v.cpu() * torch.from_numpy(G["inp"]["wv"])).sum()
    loss.backward()
    for k, prm in pol.named_parameters():
        g = prm.grad.cpu().numpy() if prm.grad is not None else np.zeros(tuple(prm.shape), dtype=np.float32)
        _check("grad " + k, g, G["grad"][k], tol=max(tol, 5e-5), abs_floor=2e-7)
    if meta["lstm_on"]:
        _check("h_pi", pol.PI.hidden_cell[0].detach().cpu().numpy(), G["out"]["h_pi"], tol=tol)
        _check("c_pi", pol.PI.hidden_cell[1].detach().cpu().numpy(), G["out"]["c_pi"], tol=tol)
