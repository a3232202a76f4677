"""Async-gather probe (csrc/s2v_probe.cu): time the ring gather of one backward round's neighbour rows on the AMS batch
for cp.async (mode 0) / cp.async.bulk per row (mode 1), with (+2) and without the own-row tile, over ring sizes and
loader-warp counts.  Checks the result against torch first.  usage: gather_probe.py [B] [layout: batched|replicated]"""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200 import _lib
from simmodel_b200.layout import _p, _stream
L = _lib.lib(); dev = torch.device('cuda'); B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kind = sys.argv[2] if len(sys.argv) > 2 else 'batched'
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes; N = B * n
if kind == 'batched':
    ei = sb.graphs.batch_edge_index(topo, B, device=dev)
    lay = sb.GraphLayout.from_batch(ei, (torch.arange(B + 1, device=dev) * n), n, validate=False)
else:
    lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B)
    ei = sb.graphs.batch_edge_index(topo, B, device=dev)
src = torch.randn(N, 64, device=dev); own = torch.randn(N, 64, device=dev); out = torch.empty(N, 64, device=dev)
g = lay.c_struct()
fn = L.s2v_debug_gather_probe
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
# backward gather: out[j] = sum over edges (i, j) of src[i]  (transposed CSR: column j lists the rows i that gather it)
ref = torch.zeros(N, 64, device=dev).index_add_(0, ei[1], src[ei[0]])
alg = 4 * (64 * (lay.num_edges / (lay.period or N) + 1) + lay.num_edges / (lay.period or N) + 1) * N
print('layout %s  N %d  E/N %.2f' % (kind, N, lay.num_edges / (lay.period or N)), flush=True)
for mode in (0, 2):
    for loaders in (4,):
        for R, cap in ((320, 128), (384, 128), (512, 256)):
            if (mode & 2) and cap + 64 > R: continue
            out.zero_()
            rc = fn(C.byref(g), _p(src), _p(own), _p(out), mode, R, cap, _stream())
            assert rc == 0, rc
            torch.cuda.synchronize()
            want = ref + own if mode & 2 else ref
            err = float((out - want).abs().max())
            for _ in range(2): fn(C.byref(g), _p(src), _p(own), _p(out), mode, R, cap, _stream())
            torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): fn(C.byref(g), _p(src), _p(own), _p(out), mode, R, cap, _stream())
            b.record(); torch.cuda.synchronize(); ms = a.elapsed_time(b) / 5
            bytes_ = alg + (256 * N if mode & 2 else 0)
            print('mode %d (%s%s) loaders %d ring %3d cap %3d: %.3f ms  %.0f GB/s  maxerr %.1e' % (
                mode, 'bulk' if mode & 1 else 'ldgsts', '+own' if mode & 2 else '', loaders, R, cap, ms, bytes_ / ms / 1e6, err), flush=True)
            prof = (C.c_longlong * 16)()
            L.s2v_debug_gather_probe_prof(prof)
            nt = max(1, prof[10])
            print('    per tile cycles: M wait %d total %d | L0 wait %d issue %d | C0 wait %d sum %d (desc %d) store %d total %d' % (
                prof[0] / nt, prof[3] / nt, prof[4] / nt, prof[6] / nt, prof[8] / nt, prof[12] / nt, prof[14] / nt, prof[13] / nt, prof[9] / nt), flush=True)
