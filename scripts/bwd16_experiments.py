"""Where does k_embed_round_bwd_tc16 spend its time?  The same launch with one piece of work switched off at a time
(libs2v_exp16.so = the library built with -DS2V_BWD16_EXPERIMENTS, scripts/build_exp16.sh; results are garbage, only
the time matters).  usage: bwd16_experiments.py [lib]"""
import ctypes as C, os, sys, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200.layout import _p, _stream
L = C.CDLL(sys.argv[1] if len(sys.argv) > 1 else 'simmodel_b200/libs2v_exp16.so'); dev = torch.device('cuda'); B = 4096
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes
lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B); N = B * n
w2 = (torch.randn(64, 64) / 8).to(dev); du = torch.randn(N, 64, device=dev); mu = torch.rand(N, 64, device=dev); out = torch.empty(N, 64, device=dev)
L.s2v_embed_backward_workspace_bytes.restype = C.c_int64
ws = torch.empty(L.s2v_embed_backward_workspace_bytes(7, 64), dtype=torch.uint8, device=dev)
g = lay.c_struct()
L.s2v_embed_round_backward.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 5 + [C.c_int64, C.c_void_p]
names = {0: 'everything on', 1: 'no reduce MMAs', 2: 'no MMAs', 4: 'no m smem stores', 8: 'no D2 read-out', 16: 'no du global stores',
         5: 'no reduce MMAs, no m stores', 13: 'no reduce MMAs / m stores / read-out', 30: 'gather + g stores + barriers only',
         32: 'no g smem stores', 64: 'no staging', 94: 'gather + g stores + barriers, no staging', 126: 'loads + barriers only'}
for x, nm in names.items():
    os.environ['S2V_B200_BWD16_EXP'] = str(x)
    def run(): assert L.s2v_embed_round_backward(C.byref(g), 7, 64, 4, _p(w2), _p(du), _p(mu), _p(out), _p(ws), ws.numel(), _stream()) == 0
    for _ in range(3): run()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(10): run()
    b.record(); torch.cuda.synchronize()
    print('%-40s %.3f ms' % (nm, a.elapsed_time(b) / 10), flush=True)
