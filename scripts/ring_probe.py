"""Forward round with the asynchronous ring gather (k_embed_round_ring) against k_embed_round_tc16: bit-identity and
time per launch over ring / chunk sizes.  usage: ring_probe.py [B] [layout: rep|bat]"""
import ctypes as C, os, sys, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200 import _lib
from simmodel_b200.layout import _p, _stream
L = _lib.lib(); dev = torch.device('cuda'); B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kind = sys.argv[2] if len(sys.argv) > 2 else 'rep'
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes; N = B * n
if kind == 'rep':
    lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B)
else:
    ei = sb.graphs.batch_edge_index(topo, B).to(dev)
    ptr = (torch.arange(B + 1, dtype=torch.int64) * n).to(dev)
    lay = sb.GraphLayout.from_batch(ei, ptr, n)
torch.manual_seed(0)
w2 = (torch.randn(64, 64) / 8).to(dev)
base = torch.randn(N, 64, device=dev); mu = torch.rand(N, 64, device=dev)
g = lay.c_struct()
def run(out): _lib.check(L.s2v_embed_round_forward(C.byref(g), 64, 3, _p(w2), _p(base), _p(mu), _p(out), _stream()), 'f')
def timed(out, reps=10):
    for _ in range(3): run(out)
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): run(out)
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
os.environ.pop('S2V_B200_FWD_RING', None)
ref = torch.empty(N, 64, device=dev); t0 = timed(ref)
print('%s B=%d rows %d | k_embed_round_tc16            %.3f ms (%.0f GB/s)' % (kind, B, N, t0, 1276.0 * N / t0 / 1e6), flush=True)
cfgs = sys.argv[3:] or ['1,0,64', '1,0,96', '1,0,128', '1,0,192', '1,0,256', '2,0,128', '2,0,192', '1,384,128', '1,480,128']
for cfg in cfgs:
    os.environ['S2V_B200_FWD_RING'] = cfg
    out = torch.full((N, 64), float('nan'), device=dev)
    try:
        t1 = timed(out)
    except Exception as e:
        print('ring %-12s FAILED %s' % (cfg, e), flush=True); break
    same = torch.equal(out, ref)
    print('ring %-12s %.3f ms (%.0f GB/s, %.2f of 6546)  bit-identical to tc16: %s  maxdiff %.3g' % (
        cfg, t1, 1276.0 * N / t1 / 1e6, 1276.0 * N / t1 / 1e6 / 6546.2, same, float((out - ref).abs().max())), flush=True)
