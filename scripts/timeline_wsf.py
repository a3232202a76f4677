"""Clock stamps of CTA 0 in the warp-specialised LAST backward stage (the last k_embed_bwd_ws launch of a backward pass)."""
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200 import _lib
_lib.LIB_PATH = 'simmodel_b200/libs2v_tl.so'
dev = torch.device('cuda'); B = 4096
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes
lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B); N = B * n
torch.manual_seed(0)
net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False).to(dev)
x = torch.rand(N, 7, device=dev)
for _ in range(2):
    net.zero_grad(); q = net.forward_graph(x, lay); q.sum().backward()
torch.cuda.synchronize()
H = C.CDLL('simmodel_b200/libs2v_tl.so')
buf = (C.c_longlong * 128)(); assert H.s2v_debug_timeline(buf) == 0
T = np.array(list(buf)).reshape(8, 16); X = T - T[0, 0]
print('P0 of tiles 0,50,..,350 (CTA 0): cycles per tile over each 50-tile span', np.diff(T[:, 11]) / 50)
print('slots: P0 start, P1 loaded+summed, P2 stage free, P3 full_g | I4 operands ready, I5 committed | F6 drain start, F7 drain end | E8 h stored, E9 staging ready, E10 dh reduced')
for k in range(6):
    print('k=%d ' % (k + 4) + ' '.join('%7d' % v for v in X[k, :11]))
print('P: period', np.diff(X[:6, 0]), ' load+sum', X[:6, 1] - X[:6, 0], ' stage wait', X[:6, 2] - X[:6, 1], ' store', X[:6, 3] - X[:6, 2])
print('E: h stored period', np.diff(X[:6, 8]), ' dh reduce', X[:6, 10] - X[:6, 9])
print('I: period', np.diff(X[:6, 4]), ' F drain', X[:6, 7] - X[:6, 6])
