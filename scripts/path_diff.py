"""Where do the CUDA-core and the tensor-core paths differ?  UTR x 5 graphs, dense random-sign loss: error of q, mu and every
parameter gradient of both paths against the fp64 oracle, relative to the tensor's scale."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simmodel_b200 as sb
from oracle import s2v_ref as R
dev = torch.device("cuda")
topo = sb.graphs.load_topology("NWB_UTR"); B = 5
gen = torch.Generator().manual_seed(17)
x = sb.graphs.synth_nfm(topo, B, gen).reshape(-1, 7)
ei = sb.graphs.batch_edge_index(topo, B); ptr = torch.arange(B + 1) * topo.num_nodes
n = B * topo.num_nodes
w = torch.randn(n, generator=gen)
P = R.make_qnet_params(7, 64, seed=5, scale=0.5)
P64 = {k: v.double().requires_grad_(True) for k, v in P.items()}
q64 = R.qnet_forward_sparse(P64, x.double(), ei, ptr, topo.num_nodes, 5, True)
(q64 * w.double()).sum().backward()
net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False)
net.load_state_dict(P); net = net.to(dev)
layout = sb.GraphLayout.from_batch(ei.to(dev), ptr.to(dev), topo.num_nodes)
res = {}
for path in ("ffma", "auto"):
    os.environ["S2V_B200_PATH"] = path
    net.zero_grad()
    q = net.forward_graph(x.to(dev), layout)
    (q * w.to(dev)).sum().backward()
    res[path] = {"q": q.detach().cpu().double()}
    for k, p_ in net.named_parameters():
        res[path][k] = p_.grad.cpu().double()
ref = {"q": q64.detach()}
ref.update({k: v.grad for k, v in P64.items()})
print("%-16s %10s %12s %12s" % ("tensor", "scale", "ffma err", "auto err"))
for k in ref:
    sc = float(ref[k].abs().max()) + 1e-30
    print("%-16s %10.3g %12.2e %12.2e" % (k, sc, float((res["ffma"][k] - ref[k]).abs().max()) / sc, float((res["auto"][k] - ref[k]).abs().max()) / sc))
