"""Stage timeline of the single-launch small-graph kernel (library built with -DS2V_SMALL_PROF)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200 import _lib
_lib.LIB_PATH = 'simmodel_b200/libs2v_sp.so'
from simmodel_b200 import ops
dev = torch.device('cuda')
H = C.CDLL('simmodel_b200/libs2v_sp.so')
for name, B in (('NWB_AMS', 1), ('MemTask', 50)):
    topo = sb.graphs.load_topology(name); N = topo.num_nodes
    torch.manual_seed(0)
    net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False).to(dev)
    x = sb.graphs.synth_nfm(topo, B, torch.Generator().manual_seed(B)).reshape(-1, 7).to(dev)
    lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), N, B)
    for _ in range(5):
        with torch.no_grad(): net.forward_graph(x, lay)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 32)(); H.s2v_debug_small_prof(buf)
    # globaltimer resolution is coarse (32 ns .. 1 us depending on the driver setting)
    t = list(buf); t0 = t[0]
    print(name, B, 'first %.1f us |' % ((t[1] - t0) / 1e3), ' '.join('sync %.1f round %.1f' % ((t[2 * r] - t[2 * r - 1]) / 1e3, (t[2 * r + 1] - t[2 * r]) / 1e3) for r in range(1, 5)),
          '| sync %.1f head/pool %.1f' % ((t[20] - t[9]) / 1e3, (t[21] - t[20]) / 1e3), ('| rows %.1f sync+mask %.1f | total %.1f us' % ((t[22] - t[21]) / 1e3, (t[23] - t[22]) / 1e3, (t[23] - t0) / 1e3)) if t[23] > t[21] else '| total %.1f us' % ((t[21] - t0) / 1e3))
