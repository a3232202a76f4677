import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200 import _lib
from simmodel_b200.layout import _p, _stream
L = C.CDLL('simmodel_b200/libs2v_tl.so'); dev = torch.device('cuda'); B = 4096
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes
lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B); N = B * n
w2 = (torch.randn(64, 64) / 8).to(dev); du = torch.randn(N, 64, device=dev); mu = torch.rand(N, 64, device=dev); out = torch.empty(N, 64, device=dev)
L.s2v_embed_backward_workspace_bytes.restype = C.c_int64
ws = torch.empty(L.s2v_embed_backward_workspace_bytes(7, 64), dtype=torch.uint8, device=dev)
g = lay.c_struct()
L.s2v_embed_round_backward.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]
for _ in range(3):
    assert L.s2v_embed_round_backward(C.byref(g), 7, 64, 2, _p(w2), _p(du), _p(mu), _p(out), _p(ws), ws.numel(), _stream()) == 0
torch.cuda.synchronize()
buf = (C.c_longlong * 128)(); assert L.s2v_debug_timeline(buf) == 0
t = np.array(list(buf)).reshape(8, 16)[:, :9]
names = ['gather+split', 'fence+sync', 'rowtmem', 'mma issue', 'mma wait', 'd1->stg', 'epilogue', 'flush']
for row in t[:5]:
    d = np.diff(row); print('tile total %6d cyc | ' % (row[-1] - row[0]) + '  '.join('%s %d' % (nm, v) for nm, v in zip(names, d)))
