"""Small-batch latency of the reference's own DQN shapes (BASELINE config 3): B=32 train step, B=1 acting,
eager vs captured in a CUDA graph."""
import sys, time, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
dev = torch.device('cuda'); topo = sb.graphs.load_topology('NWB_AMS'); N = topo.num_nodes
torch.manual_seed(0)
net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False).to(dev)
opt = torch.optim.Adam(net.parameters(), lr=1e-4, capturable=True)
loss_fn = torch.nn.SmoothL1Loss()
def bench(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for B in (32, 1):
    gen = torch.Generator().manual_seed(B)
    x = sb.graphs.synth_nfm(topo, B, gen).reshape(-1, 7).to(dev)
    lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), N, B)
    sel = (torch.randint(0, N, (B,), generator=gen) + torch.arange(B) * N).to(dev); tg = torch.randn(B, generator=gen).to(dev)
    def train_step():
        opt.zero_grad(set_to_none=False)
        q = net.forward_graph(x, lay); loss = loss_fn(q[sel], tg); loss.backward(); opt.step(); return loss
    def act():
        with torch.no_grad(): return net.forward_graph(x, lay)
    e_train, e_act = bench(train_step), bench(act)
    # CUDA graphs
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): train_step()
    torch.cuda.current_stream().wait_stream(s)
    gt = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gt): l_static = train_step()
    ga = torch.cuda.CUDAGraph()
    with torch.cuda.graph(ga): q_static = act()
    g_train, g_act = bench(gt.replay), bench(ga.replay)
    print('B=%d: train step eager %.3f ms  graph %.3f ms | forward(no grad) eager %.3f ms graph %.3f ms  -> %.0f graphs/s train (graph)' % (B, e_train, g_train, e_act, g_act, B / g_train * 1e3))
