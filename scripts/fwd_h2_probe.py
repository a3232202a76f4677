"""Forward round: accuracy of one round against fp64 (B = 64, both layouts) and time at B (default 4096) for the kernel
selected by S2V_B200_FWD_ROUND (tc16 | h2[,CTAs per SM]).  usage: fwd_h2_probe.py [B]"""
import ctypes as C, os, sys, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200 import _lib
from simmodel_b200.layout import _p, _stream
L = _lib.lib(); dev = torch.device('cuda'); B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes
tag = os.environ.get('S2V_B200_FWD_ROUND', '(default)')
for layout in ('replicated', 'batched'):
    Bs = 64; N = Bs * n
    bei = sb.graphs.batch_edge_index(topo, Bs).to(dev)
    lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, Bs) if layout == 'replicated' else \
        sb.GraphLayout.from_batch(bei, (torch.arange(Bs + 1) * n).to(dev), n)
    gen = torch.Generator().manual_seed(7)
    w2 = (torch.randn(64, 64, generator=gen) / 8).to(dev)
    base = torch.randn(N, 64, generator=gen).to(dev)
    mu = (torch.randn(N, 64, generator=gen).clamp(min=0) * torch.rand(N, 1, generator=gen) * 50).to(dev)
    agg = torch.zeros(N, 64, dtype=torch.float64, device=dev).index_add_(0, bei[0], mu.double()[bei[1]])
    pre = base.double() + agg @ w2.double().t()
    ref = pre.clamp(min=0)
    g = lay.c_struct()
    for flag in (1, 3):
        out = torch.full((N, 64), float('nan'), device=dev)
        _lib.check(L.s2v_embed_round_forward(C.byref(g), 64, flag, _p(w2), _p(base), _p(mu), _p(out), _stream()), 'f')
        torch.cuda.synchronize()
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        flips = int(((out > 0) != (pre > 0)).sum())
        print('%-10s %-10s flag %d  mu_next err %.2e  relu decisions differing from fp64: %d of %d' % (tag, layout, flag, err, flips, out.numel()), flush=True)
N = B * n
lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B)
w2 = (torch.randn(64, 64) / 8).to(dev); base = torch.randn(N, 64, device=dev); mu = torch.rand(N, 64, device=dev); out = torch.empty(N, 64, device=dev)
g = lay.c_struct()
def run(): _lib.check(L.s2v_embed_round_forward(C.byref(g), 64, 3, _p(w2), _p(base), _p(mu), _p(out), _stream()), 'f')
for rep in range(2):
    for _ in range(3): run()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(20): run()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print('%-10s forward round %.3f ms  (%.0f GB/s algorithmic, %.2f of 6546)' % (tag, ms, 1276.0 * N / ms / 1e6, 1276.0 * N / ms / 1e6 / 6546.2), flush=True)
