import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200.layout import _p, _stream
L = C.CDLL('simmodel_b200/libs2v_tl.so'); dev = torch.device('cuda'); B = 4096
FLAG = int(sys.argv[1]) if len(sys.argv) > 1 else 3
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes
lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B); N = B * n
w2 = (torch.randn(64, 64) / 8).to(dev); du = torch.randn(N, 64, device=dev); mu = torch.rand(N, 64, device=dev); out = torch.empty(N, 64, device=dev)
L.s2v_embed_backward_workspace_bytes.restype = C.c_int64
ws = torch.empty(L.s2v_embed_backward_workspace_bytes(7, 64), dtype=torch.uint8, device=dev)
g = lay.c_struct()
L.s2v_embed_round_backward.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]
for _ in range(3):
    assert L.s2v_embed_round_backward(C.byref(g), 7, 64, FLAG, _p(w2), _p(du), _p(mu), _p(out), _p(ws), ws.numel(), _stream()) == 0
torch.cuda.synchronize()
buf = (C.c_longlong * 128)(); assert L.s2v_debug_timeline(buf) == 0
T = np.array(list(buf)).reshape(8, 16); t = T[:, :12]; X = T - T[0, 0]
print('P0 of tiles 0,50,..,350 (CTA 0): cycles per tile over each 50-tile span', np.diff(T[:, 11]) / 50)
t0 = t[0, 0]
print('flag', FLAG, 'slots: P0 gather issue, P1 gathered, P2 stage free, P3 full_g arrive | I4 operands ready, I5 committed | F6 drain start, F7 drain end | E8 m stored, E9 staging ready, E10 du stored')
for k, row in enumerate(t[:6]):
    print('k=%d ' % (k + 4) + ' '.join('%7d' % (v - t0) for v in row[:11]))
print('P: period', np.diff(t[:6, 0]), ' gather', (t[:6, 1] - t[:6, 0]), ' stage wait', (t[:6, 2] - t[:6, 1]), ' store', (t[:6, 3] - t[:6, 2]))
print('I: issue', (t[:6, 5] - t[:6, 4]), ' F: drain', (t[:6, 7] - t[:6, 6]), ' | E: du store', (t[:6, 10] - t[:6, 9]))
print('issuer warp, per loop iteration k (tile k issue, tile k-1 drain):')
for k in range(1, 6):
    print('  k=%d: operands ready %6d | transform issued +%5d | mma_done(k-1) +%5d | staging free +%5d | drained +%5d | full_m/d2_free ok +%5d | reduce issued, committed +%5d | next operands ready +%5d'
          % (k + 4, X[k, 4], X[k, 12] - X[k, 4], X[k - 1, 13] - X[k, 12], X[k - 1, 6] - X[k - 1, 13], X[k - 1, 7] - X[k - 1, 6], X[k, 14] - X[k - 1, 7], X[k, 5] - X[k, 14], X[k + 1, 4] - X[k, 5] if k < 5 else 0))
