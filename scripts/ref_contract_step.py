"""The reference's own call, unchanged: QFunction.batch_update(states_tsrs, Ws, pyg_data, actions, targets) with
lists of CPU tensors and dense padded 975x975 adjacencies (comb_opt.py:334-366), and get_best_action (309-331)."""
import sys, time, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
dev = torch.device('cuda'); topo = sb.graphs.load_topology('NWB_AMS'); N = topo.num_nodes; B = 32
torch.manual_seed(0)
net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False).to(dev)
opt = torch.optim.Adam(net.parameters(), lr=1e-3); sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.99999)
qf = sb.QFunction(net, opt, sched)
gen = torch.Generator().manual_seed(1)
W = torch.zeros(N, N); W[topo.edge_index[0], topo.edge_index[1]] = 1.0
xs = list(sb.graphs.synth_nfm(topo, B, gen)); Ws = [W.clone() for _ in range(B)]
datas = [sb.Data(torch.ones(N), [0, 0], num_nodes=N) for _ in range(B)]
actions = torch.randint(0, N, (B,), generator=gen).tolist(); targets = torch.randn(B, generator=gen).tolist()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print('batch_update (B=32, dense W lists from host): %.2f ms' % t(lambda: qf.batch_update(xs, Ws, datas, actions, targets)))
nb = topo.neighbors(5)
print('get_best_action (B=1, dense W from host):      %.2f ms' % t(lambda: qf.get_best_action(xs[0], W, datas[0], nb)))
Wd = W.to(dev); xd = xs[0].to(dev)
print('get_best_action (B=1, tensors on device):      %.2f ms' % t(lambda: qf.get_best_action(xd, Wd, datas[0], nb)))
ev = sb.Data(xs[0], topo.edge_index)   # evaluation path: pyg_data carries a real edge_index (rl_utils.py:26-29)
print('get_best_action (B=1, real edge_index in pyg): %.2f ms' % t(lambda: qf.get_best_action(xd, Wd, ev, nb)))
