"""GPU parity of the SB3 consumer of the s2v body against goldens produced by the UNMODIFIED
modules/ppo/models_sb3_s2v.py classes (oracle/gen_golden_sb3.py): serialised features, actor logits over all padded
nodes, critic value (max over reachable nodes), every parameter gradient, and the greedy masked action."""
import types

import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden
from test_gpu_parity import _check

pytestmark = pytest.mark.gpu


def _build(meta, G, dev):
    import simmodel_b200 as sb
    p, n_pad = meta["emb_dim"], meta["n_pad"]
    space = types.SimpleNamespace(spaces={"W": types.SimpleNamespace(shape=(n_pad, n_pad))})
    mods = {"extractor": sb.sb3.Struc2VecExtractor(space, p, meta["T"], 7), "acnet": sb.sb3.s2v_ACNetwork(p, 1, 1, p),
            "pnet": torch.nn.Linear(1, 1, True), "vnet": torch.nn.Linear(1, 1, True)}
    for g, mod in mods.items():
        sd = {k[len(g) + 1:]: torch.from_numpy(v) for k, v in G["P"].items() if k.startswith(g + ".")}
        mod.load_state_dict(sd, strict=True)
        mod.to(dev)
    return mods, space


@pytest.mark.parametrize("case", golden_cases("sb3"))
def test_sb3_golden(case, cuda_device):
    import simmodel_b200 as sb
    meta, G = load_golden(case)
    dev = cuda_device
    mods, space = _build(meta, G, dev)
    obs = {k: torch.from_numpy(G["inp"][k]) for k in ("nfm", "W", "reachable_nodes", "num_nodes")}   # host tensors, as SB3 passes
    feats = mods["extractor"](obs)
    a, b = mods["acnet"](feats)
    logits, value = mods["pnet"](a), mods["vnet"](b)
    assert feats.shape == G["out"]["features"].shape and logits.shape == G["out"]["logits"].shape and value.shape == G["out"]["value"].shape
    _check("features", feats.detach().cpu().numpy(), G["out"]["features"], G["out"]["features_f64"])
    _check("logits", logits.detach().cpu().numpy(), G["out"]["logits"], G["out"]["logits_f64"])
    _check("value", value.detach().cpu().numpy(), G["out"]["value"], G["out"]["value_f64"])
    # the reachable flags and node counts travel through the serialisation unchanged (bit-exact)
    assert np.array_equal(feats.detach().cpu().numpy()[:, :-1, -1:], G["inp"]["reachable_nodes"])
    assert np.array_equal(feats.detach().cpu().numpy()[:, -1, -1], G["inp"]["num_nodes"][:, 0])
    wl, wv = torch.from_numpy(G["inp"]["wl"]).to(dev), torch.from_numpy(G["inp"]["wv"]).to(dev)
    ((logits * wl).sum() + (value * wv).sum()).backward()
    for g, mod in mods.items():
        for k, prm in mod.named_parameters():
            _check("grad %s.%s" % (g, k), prm.grad.cpu().numpy(), G["grad"][g + "." + k], G["grad64"][g + "." + k], abs_floor=1e-7)
    if meta["greedy_action"] >= 0:
        env = types.SimpleNamespace(observation_space=space, F=7)
        trained = types.SimpleNamespace(features_extractor=mods["extractor"], mlp_extractor=mods["acnet"],
                                        action_net=mods["pnet"], value_net=mods["vnet"])
        dp = sb.sb3.DeployablePPOPolicy(env, trained, device=dev)
        o0 = {k: v[:1] for k, v in obs.items()}
        action, _ = dp.predict(o0, deterministic=True, action_masks=G["inp"]["reachable_nodes"][0])
        assert int(action) == meta["greedy_action"]
        dist = dp.get_distribution(o0)
        ref_logits = np.where(G["inp"]["reachable_nodes"][0, :, 0] != 0, G["out"]["logits"][0, :, 0], np.float32(-1e20))
        ref = torch.distributions.Categorical(logits=torch.from_numpy(ref_logits)).probs.numpy()
        _check("distribution", dist.probs.detach().cpu().numpy().reshape(-1), ref, tol=2e-5)
        _check("predict_values", dp.predict_values(o0).detach().cpu().numpy(), G["out"]["value"][:1], tol=1e-5)


def test_sb3_no_cpu_path():
    import simmodel_b200 as sb
    ex = sb.sb3.Struc2VecExtractor(5, 16, 2, 7)
    with pytest.raises(sb.S2VError):
        ex.get_embeddings(torch.zeros(1, 5, 7), torch.zeros(1, 5, 5))
