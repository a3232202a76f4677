"""Time one propagation round (forward / backward) per kernel variant on the AMS batch: flags 1 ffma, 2 tc,
3 tcws (warp-specialised backward).  usage: round_probe.py [B] [flags...]"""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200 import _lib
from simmodel_b200.layout import _p, _stream
import os
if os.environ.get('S2V_PROBE_LIB'): _lib.LIB_PATH = os.environ['S2V_PROBE_LIB']
L = _lib.lib(); dev = torch.device('cuda'); B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
flags = [int(a) for a in sys.argv[2:]] or [1, 2, 3]
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes
lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B); N = B * n
w2 = (torch.randn(64, 64) / 8).to(dev)
base = torch.randn(N, 64, device=dev); mu = torch.rand(N, 64, device=dev); out = torch.empty(N, 64, device=dev)
du = torch.randn(N, 64, device=dev)
ws = torch.empty(L.s2v_embed_backward_workspace_bytes(7, 64), dtype=torch.uint8, device=dev)
g = lay.c_struct()
names = {1: 'ffma', 2: 'tc', 3: 'tcws', 4: 'tcws2'}
for flag in flags:
    def fwd(): _lib.check(L.s2v_embed_round_forward(C.byref(g), 64, flag, _p(w2), _p(base), _p(mu), _p(out), _stream()), 'f')
    def bwd(): _lib.check(L.s2v_embed_round_backward(C.byref(g), 7, 64, flag, _p(w2), _p(du), _p(mu), _p(out), _p(ws), ws.numel(), _stream()), 'b')
    res = []
    for fn in (fwd, bwd):
        for _ in range(3): fn()
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): fn()
        b.record(); torch.cuda.synchronize(); res.append(a.elapsed_time(b) / 10)
    gbs = [1276.0 * N / (t * 1e-3) / 1e9 for t in res]
    print('%-6s fwd %.3f ms (%.0f GB/s)  bwd %.3f ms (%.0f GB/s, %.2f of 6546)' % (names[flag], res[0], gbs[0], res[1], gbs[1], gbs[1] / 6546.2), flush=True)
