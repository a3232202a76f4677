"""Generate tests/golden/sb3_*.npz by running the UNMODIFIED reference classes of modules/ppo/models_sb3_s2v.py
(TEST INFRASTRUCTURE ONLY; build container only:  python -m oracle.gen_golden_sb3).

Struc2VecExtractor (32-148), s2v_ACNetwork (150-238) and DeployablePPOPolicy (361-420) import from /root/reference as
they are; the absent packages (gym, stable_baselines3, sb3_contrib, torch_scatter) are stand-ins: BaseFeaturesExtractor
is an nn.Module that stores features_dim (what SB3's does), scatter_mean is oracle/s2v_ref.scatter_mean.  The same
classes are run a second time in float64 (torch default dtype switched) for adjudication.
"""
from __future__ import annotations

import importlib
import types

import numpy as np
import torch

from oracle import ref_import
from oracle.gen_golden import _np, _random_digraph, _save, _scale_of
from simmodel_b200 import graphs


def load_sb3_reference():
    ref_import._ABSENT_ROOTS.add("simdata_utils")          # `import simdata_utils as su` (unused on this path)
    co, _ = ref_import.load_reference()
    import gym.spaces  # noqa: F401  (fabricated; makes gym.spaces.Dict resolvable in annotations)
    import stable_baselines3.common.torch_layers as tl

    class BaseFeaturesExtractor(torch.nn.Module):
        def __init__(self, observation_space, features_dim):
            super().__init__()
            self._features_dim = features_dim
    tl.BaseFeaturesExtractor = BaseFeaturesExtractor
    m = importlib.import_module("modules.ppo.models_sb3_s2v")
    m.scatter_mean = co.scatter_mean
    m.device = "cpu"
    return m


def _obs(sizes, eis, xs, n_pad, reach):
    B = len(sizes)
    nfm = torch.zeros(B, n_pad, graphs.NODE_DIM)
    W = torch.zeros(B, n_pad, n_pad)
    r = torch.zeros(B, n_pad, 1)
    for b in range(B):
        nfm[b, :sizes[b]] = xs[b]
        W[b, eis[b][0], eis[b][1]] = 1.0
        r[b, reach[b], 0] = 1.0
    return {"nfm": nfm, "W": W, "reachable_nodes": r, "num_nodes": torch.tensor([[float(s)] for s in sizes])}


def case(m, name, sizes, eis, xs, n_pad, p, T, seed):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    space = types.SimpleNamespace(spaces={"W": types.SimpleNamespace(shape=(n_pad, n_pad))})
    ex = m.Struc2VecExtractor(space, p, T, graphs.NODE_DIM)
    ac = m.s2v_ACNetwork(p, 1, 1, p)
    pnet, vnet = torch.nn.Linear(1, 1, True), torch.nn.Linear(1, 1, True)
    reach = []
    for b, s in enumerate(sizes):
        cur = int(torch.randint(0, s, (1,), generator=gen))
        nb = sorted(set(eis[b][1][eis[b][0] == cur].tolist())) or [cur]
        reach.append(nb)
    obs = _obs(sizes, eis, xs, n_pad, reach)
    mods = {"extractor": ex, "acnet": ac, "pnet": pnet, "vnet": vnet}
    wl = torch.randn(len(sizes), n_pad, 1, generator=gen)
    wv = torch.randn(len(sizes), 1, generator=gen)

    def run(dtype):
        o = {k: v.to(dtype) for k, v in obs.items()}
        feats = ex(o)
        a, b = ac(feats)
        logits, value = pnet(a), vnet(b)
        for mod in mods.values():
            mod.zero_grad()
        ((logits * wl.to(dtype)).sum() + (value * wv.to(dtype)).sum()).backward()
        grads = {g + "." + k: prm.grad.detach().clone() for g, mod in mods.items() for k, prm in mod.named_parameters()}
        return feats.detach(), logits.detach(), value.detach(), grads
    feats, logits, value, grads = run(torch.float32)
    # greedy action of graph 0 alone through the reference's DeployablePPOPolicy (B = 1, unpadded is not possible here:
    # the extractor is built for n_pad; the policy is "invariant to the number of nodes" through padding)
    env = types.SimpleNamespace(observation_space=space, F=graphs.NODE_DIM)
    trained = types.SimpleNamespace(features_extractor=ex, mlp_extractor=ac, action_net=pnet, value_net=vnet)
    if p == 64 and T == 5:
        dp = m.DeployablePPOPolicy(env, trained)
        o0 = {k: v[:1] for k, v in obs.items()}
        action, _ = dp.predict(o0, deterministic=True, action_masks=obs["reachable_nodes"][0].numpy())
        action = int(action)
    else:
        action = -1
    P = {g + "." + k: v.detach().clone() for g, mod in mods.items() for k, v in mod.state_dict().items()}
    # fp64 adjudication: the same classes in double
    torch.set_default_dtype(torch.float64)
    try:
        for mod in mods.values():
            mod.double()
        f64, l64, v64, g64 = run(torch.float64)
    finally:
        torch.set_default_dtype(torch.float32)
        for mod in mods.values():
            mod.float()
    own = max(float((grads[k].double() - g64[k]).abs().max()) / _scale_of(g64[k]) for k in grads)
    print("  %-26s logits scale %.3g value scale %.3g | reference's own fp32 error vs fp64: features %.1e logits %.1e value %.1e grads %.1e"
          % (name, _scale_of(logits), _scale_of(value), float((feats.double() - f64).abs().max()) / _scale_of(f64),
             float((logits.double() - l64).abs().max()) / _scale_of(l64), float((value.double() - v64).abs().max()) / _scale_of(v64), own))
    meta = dict(kind="sb3", sizes=sizes, n_pad=n_pad, emb_dim=p, T=T, seed=seed, greedy_action=action)
    _save(name, meta, P={k: _np(v) for k, v in P.items()},
          inp={k: _np(v) for k, v in obs.items()} | {"wl": _np(wl), "wv": _np(wv)},
          out=dict(features=_np(feats), logits=_np(logits), value=_np(value), features_f64=_np(f64), logits_f64=_np(l64),
                   value_f64=_np(v64)),
          grad={k: _np(v) for k, v in grads.items()}, grad64={k: _np(v) for k, v in g64.items()})


def main():
    m = load_sb3_reference()
    gen = torch.Generator().manual_seed(777)
    print("SB3 s2v cases (Struc2VecExtractor + s2v_ACNetwork + DeployablePPOPolicy)")
    m3, m5 = graphs.load_topology("Manhattan3"), graphs.load_topology("Manhattan5")
    xs = [graphs.synth_nfm(m3, 1, gen)[0], graphs.synth_nfm(m5, 1, gen)[0], graphs.synth_nfm(m3, 1, gen)[0]]
    case(m, "sb3_m3m5mix_pad25", [9, 25, 9], [m3.edge_index, m5.edge_index, m3.edge_index], xs, 25, 64, 5, seed=40)
    sizes = [5, 9, 1, 12, 7]
    eis = [_random_digraph(s, 2.5, gen) for s in sizes]
    xs = [torch.randn(s, graphs.NODE_DIM, generator=gen).abs() for s in sizes]
    case(m, "sb3_directed_pad14_p24_T3", sizes, eis, xs, 14, 24, 3, seed=41)


if __name__ == "__main__":
    main()
