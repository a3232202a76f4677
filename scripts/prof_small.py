import sys, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
dev = torch.device('cuda')
for name, B in (('Manhattan5', 32), ('MemTask', 50)):
    topo = sb.graphs.load_topology(name); N = topo.num_nodes
    torch.manual_seed(0)
    net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False).to(dev)
    x = sb.graphs.synth_nfm(topo, B, torch.Generator().manual_seed(B)).reshape(-1, 7).to(dev)
    lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), N, B)
    rptr = (torch.arange(B + 1, dtype=torch.int32) * 3).to(dev); ridx = torch.tensor([0, 1, 2] * B, dtype=torch.int32, device=dev)
    mask = (torch.rand(B * N) < 0.4).to(dev)
    for _ in range(3):
        with torch.no_grad():
            net.greedy_actions_graph(x, lay, rptr, ridx)
            mu = net.embed_graph(x, lay)
            sb.ops.s2v_head_softmax(lay, mu, net.head_weights(), mask)
    torch.cuda.synchronize()
print('ok')
