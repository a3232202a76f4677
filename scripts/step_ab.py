"""A/B timing of whole-library builds on the benchmarked step (AMS, B graphs, DQN loss: one node per graph):
step_ab.py B lib1.so lib2.so ...   Each library runs in its own process (the product path binds one library)."""
import os, subprocess, sys
if len(sys.argv) > 2 and sys.argv[1] != '--child':
    B = sys.argv[1]
    for rep in range(2):
        for lib in sys.argv[2:]:
            subprocess.run([sys.executable, __file__, '--child', B, lib], check=True)
    sys.exit(0)
import torch
sys.path.insert(0, '.')
B, lib = int(sys.argv[2]), sys.argv[3]
from simmodel_b200 import _lib
_lib.LIB_PATH = os.path.abspath(lib)
import simmodel_b200 as sb
dev = torch.device('cuda')
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes
lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B); N = B * n
torch.manual_seed(0)
net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False).to(dev)
x = torch.rand(N, 7, device=dev)
sel = torch.arange(B, device=dev) * n + torch.randint(0, n, (B,), device=dev)
tgt = torch.randn(B, device=dev)
def step():
    net.zero_grad(set_to_none=True)
    q = net.forward_graph(x, lay)
    torch.nn.functional.smooth_l1_loss(q[sel], tgt).backward()
for _ in range(3): step()
torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
for _ in range(10): step()
b.record(); torch.cuda.synchronize()
print('%-36s fwd+bwd %.3f ms / step' % (lib, a.elapsed_time(b) / 10), flush=True)
