"""Producer-warp cycle counters of the last backward stage (library built with -DS2V_WSF_PROF): wsf_prof.py lib.so"""
import ctypes as C, os, sys, torch
sys.path.insert(0, '.')
from simmodel_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
import simmodel_b200 as sb
dev = torch.device('cuda'); B = 4096
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes
lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B); N = B * n
torch.manual_seed(0)
net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False).to(dev)
x = torch.rand(N, 7, device=dev)
for _ in range(3):
    net.zero_grad(); q = net.forward_graph(x, lay); q[::n].sum().backward()
torch.cuda.synchronize()
H = C.CDLL(_lib.LIB_PATH); buf = (C.c_longlong * 16)(); assert H.s2v_debug_wsf_prof(buf) == 0
v = list(buf); nt = max(v[6], 1)
print('%s: P warp 0, cycles per tile (%d tiles): [0] %d | [1] %d | [2] %d | [3] %d | [4] %d | total %d   (PROF=1: wait landed, read+add, re-arm, wait stage, split+store; PROF=3: [1] before the plane loop, [2] plane loop, [3] column sums)'
      % (sys.argv[1], nt, v[0] / nt, v[1] / nt, v[2] / nt, v[3] / nt, v[4] / nt, v[5] / nt))
