"""A/B timing of forward-round builds: fwd_ab.py lib1.so lib2.so ... (each timed 3 x 20 launches, interleaved)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200.layout import _p, _stream
dev = torch.device('cuda'); B = 4096
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes
lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B); N = B * n
w2 = (torch.randn(64, 64) / 8).to(dev); base = torch.randn(N, 64, device=dev); mu = torch.rand(N, 64, device=dev)
outs = []
libs = []
for path in sys.argv[1:]:
    L = C.CDLL(path)
    L.s2v_embed_round_forward.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 5
    libs.append((path, L)); outs.append(torch.empty(N, 64, device=dev))
g = lay.c_struct()
for rep in range(3):
    for (path, L), out in zip(libs, outs):
        def run(): assert L.s2v_embed_round_forward(C.byref(g), 64, 3, _p(w2), _p(base), _p(mu), _p(out), _stream()) == 0
        for _ in range(3): run()
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
        for _ in range(20): run()
        b.record(); torch.cuda.synchronize()
        print('%-32s %.3f ms   same bits as the first library: %s' % (path, a.elapsed_time(b) / 20, torch.equal(out, outs[0])), flush=True)
