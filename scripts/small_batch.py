"""Latency of the reference's small call shapes through the public API (eager and CUDA-graph replay):
B = 1 acting forward on AMS, Manhattan5 B = 32 train step, MemTask rollout.  usage: small_batch.py"""
import sys, time, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200 import ops
dev = torch.device('cuda')


def wall(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6


def graphed(fn):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    return g.replay


for name, B in (('NWB_AMS', 1), ('NWB_AMS', 32), ('Manhattan5', 32), ('MemTask', 50)):
    topo = sb.graphs.load_topology(name); N = topo.num_nodes
    torch.manual_seed(0)
    net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False).to(dev)
    gen = torch.Generator().manual_seed(B)
    x = sb.graphs.synth_nfm(topo, B, gen).reshape(-1, 7).to(dev)
    lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), N, B)
    rptr = (torch.arange(B + 1, dtype=torch.int32) * 3).to(dev)
    ridx = torch.tensor([0, 1, 2] * B, dtype=torch.int32, device=dev)

    def act():
        with torch.no_grad():
            return net.greedy_actions_graph(x, lay, rptr, ridx)

    def fwd():
        with torch.no_grad():
            return net.forward_graph(x, lay)
    print('%-10s B=%-3d rows %-6d | greedy action: eager %.1f us, graph %.1f us | q forward: eager %.1f us, graph %.1f us' % (
        name, B, B * N, wall(act), wall(graphed(act)), wall(fwd), wall(graphed(fwd))), flush=True)
