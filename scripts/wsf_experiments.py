"""Where does the warp-specialised LAST backward stage spend its time?  Whole fwd+bwd step (AMS, B graphs) with one
piece of the final kernel switched off at a time (libs2v_exp.so, -DS2V_WS_EXPERIMENTS; only time differences matter)."""
import os, sys, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200 import _lib
_lib.LIB_PATH = 'simmodel_b200/libs2v_exp.so'
dev = torch.device('cuda'); B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes
lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B); N = B * n
torch.manual_seed(0)
net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False).to(dev)
x = torch.rand(N, 7, device=dev)
names = {0: 'everything on', 3: 'no MMAs', 4: 'no dh reductions (E)', 8: 'no D smem stores (P)', 16: 'no h smem stores (E)',
         32: 'no staging writes (F)', 63: 'loads + barriers only'}
base = None
for xf, nm in names.items():
    os.environ['S2V_WS_EXP_FINAL'] = str(xf)
    def step():
        net.zero_grad(set_to_none=True); q = net.forward_graph(x, lay); q.sum().backward()
    for _ in range(2): step()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(5): step()
    b.record(); torch.cuda.synchronize(); ms = a.elapsed_time(b) / 5
    base = base or ms
    print('%-26s step %.3f ms  (final stage %+.3f ms)' % (nm, ms, ms - base), flush=True)
