"""ReLU-kink check: embeddings of every round on the CUDA-core and the tensor-core paths -- how many activations have a
different sign pattern (mu > 0), how large they are, and the per-round max |difference| of mu relative to its scale."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simmodel_b200 as sb
from simmodel_b200 import ops
from oracle import s2v_ref as R
dev = torch.device("cuda")
topo = sb.graphs.load_topology("NWB_UTR"); B = 5
gen = torch.Generator().manual_seed(17)
x = sb.graphs.synth_nfm(topo, B, gen).reshape(-1, 7).to(dev)
ei = sb.graphs.batch_edge_index(topo, B); ptr = torch.arange(B + 1) * topo.num_nodes
P = R.make_qnet_params(7, 64, seed=5, scale=0.5)
net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False)
net.load_state_dict(P); net = net.to(dev)
layout = sb.GraphLayout.from_batch(ei.to(dev), ptr.to(dev), topo.num_nodes)
w = [t.detach() for t in net.embed_weights()]
mus = {}
for path in ("ffma", "auto"):
    os.environ["S2V_B200_PATH"] = path
    base, m = ops._embed_forward(layout, 5, x, w)
    mus[path] = m.clone()
for t in range(5):
    a, b = mus["ffma"][t], mus["auto"][t]
    flip = (a > 0) != (b > 0)
    print("round %d: scale %.3g  max|diff|/scale %.2e  sign mismatches %d of %d  (largest value among them %.2e)" % (
        t + 1, float(a.abs().max()), float((a - b).abs().max() / a.abs().max()), int(flip.sum()), a.numel(),
        float(torch.maximum(a, b)[flip].max()) if bool(flip.any()) else 0.0))
