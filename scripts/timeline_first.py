import ctypes as C, sys, numpy as np, torch, os
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200 import _lib
_lib.LIB_PATH = 'simmodel_b200/libs2v_tl.so'
L = _lib.lib(); dev = torch.device('cuda'); B = 4096
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes
lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B); N = B * n
torch.manual_seed(0)
net = sb.QNet({"emb_dim": 64, "emb_iter_T": 1, "norm_agg": True, "node_dim": 7}, verbose=False).to(dev)
x = torch.rand(N, 7, device=dev)
with torch.no_grad():
    for _ in range(3): net.embed_graph(x, lay)
torch.cuda.synchronize()
H = C.CDLL('simmodel_b200/libs2v_tl.so')
buf = (C.c_longlong * 128)(); assert H.s2v_debug_timeline(buf) == 0
t = np.array(list(buf)).reshape(8, 16)[:, :9]
names = ['x load+sync', 'h+split', 'fence+sync', 'mma issue', 'mma wait', 'd1->stg+sync', 'epilogue stores', 'sync']
for row in t[:5]:
    d = np.diff(row); print('tile total %6d cyc | ' % (row[-1] - row[0]) + '  '.join('%s %d' % (nm, v) for nm, v in zip(names, d)))
print('tile-to-tile', np.diff(t[:5, 0]))
