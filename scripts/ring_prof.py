"""Per-role wait / work cycles of CTA 0 in k_embed_round_ring (library built with -DS2V_RING_PROF).
usage: ring_prof.py cfg [cfg ...]   (cfg = stages,slots,cap)"""
import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, '.')
from simmodel_b200 import _lib
_lib.LIB_PATH = os.path.abspath('simmodel_b200/libs2v_rprof.so')
import simmodel_b200 as sb
from simmodel_b200.layout import _p, _stream
L = _lib.lib(); dev = torch.device('cuda'); B = 4096
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes; N = B * n
lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B)
w2 = (torch.randn(64, 64) / 8).to(dev); base = torch.randn(N, 64, device=dev); mu = torch.rand(N, 64, device=dev)
out = torch.empty(N, 64, device=dev); g = lay.c_struct()
H = C.CDLL(_lib.LIB_PATH)
names = ['M wait ring', 'M total', 'L wait meta', 'L issue', 'C wait full', 'C sum', 'C wait a_free', 'C split+store', 'C total',
         'X wait', 'X issue', 'E wait d_full', 'E work', 'tiles', 'E tmem->stg', 'E bar1', 'E stg->global']
for cfg in sys.argv[1:]:
    os.environ['S2V_B200_FWD_RING'] = cfg
    for _ in range(3):
        _lib.check(L.s2v_embed_round_forward(C.byref(g), 64, 3, _p(w2), _p(base), _p(mu), _p(out), _stream()), 'f')
    torch.cuda.synchronize()
    buf = (C.c_longlong * 32)(); assert H.s2v_debug_ring_prof(buf) == 0
    v = np.array(list(buf)); tiles = max(int(v[13]), 1)
    print('cfg %s: cycles per tile (CTA 0, %d tiles): ' % (cfg, tiles) + ', '.join('%s %d' % (nm, v[i] / tiles) for i, nm in enumerate(names) if i != 13), flush=True)
