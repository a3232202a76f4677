"""The benchmarked train step (AMS, B graphs, DQN loss, Adam) eager against the same step replayed from a CUDA graph.
usage: step_graph_ab.py [B]"""
import sys, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device('cuda')
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes
ei = sb.graphs.batch_edge_index(topo, B).to(dev); ptr = (torch.arange(B + 1) * n).to(dev)
lay = sb.GraphLayout.from_batch(ei, ptr, n, validate=False); N = B * n
del ei
torch.manual_seed(0)
net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False).to(dev)
opt = torch.optim.Adam(net.parameters(), lr=1e-4, capturable=True)
x = torch.rand(N, 7, device=dev)
sel = torch.arange(B, device=dev) * n + torch.randint(0, n, (B,), device=dev)
tgt = torch.randn(B, device=dev)
loss_fn = torch.nn.SmoothL1Loss()
def step():
    opt.zero_grad(set_to_none=True)
    q = net.forward_graph(x, lay)
    loss = loss_fn(q[sel], tgt)
    loss.backward()
    opt.step()
    return loss
def timed(fn, k=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(k): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / k
print('eager        %.3f ms / step' % timed(step), flush=True)
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(side)
g = torch.cuda.CUDAGraph()
opt.zero_grad(set_to_none=True)
with torch.cuda.graph(g):
    step()
print('CUDA graph   %.3f ms / step' % timed(g.replay), flush=True)
print('eager again  %.3f ms / step' % timed(step), flush=True)
