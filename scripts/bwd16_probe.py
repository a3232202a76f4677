"""Backward round: the forward-shaped 4-warp kernel (flags 4, k_embed_round_bwd_tc16) against the warp-specialised one
(flags 3) -- one round against fp64 at B = 64 (du_prev, dtheta2, zero pattern), then timing at B (default 4096).
usage: bwd16_probe.py [B]      env S2V_B200_BWD16_GROUP=<tiles chained in D2>"""
import ctypes as C, os, sys, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200 import _lib
from simmodel_b200.layout import _p, _stream
L = _lib.lib(); dev = torch.device('cuda'); B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes
stride = 2 * 64 * 64 + 4 * 64 + 64 * 7
print('S2V_B200_BWD16_GROUP', os.environ.get('S2V_B200_BWD16_GROUP', '(default)'))


def one(lay, N, flag, w2, du, mu):
    out = torch.full((N, 64), float('nan'), device=dev)
    ws = torch.zeros(L.s2v_embed_backward_workspace_bytes(7, 64), dtype=torch.uint8, device=dev)
    g = lay.c_struct()
    _lib.check(L.s2v_embed_round_backward(C.byref(g), 7, 64, flag, _p(w2), _p(du), _p(mu), _p(out), _p(ws), ws.numel(), _stream()), 'b')
    torch.cuda.synchronize()
    part = ws.view(torch.float32)[: 1024 * stride].view(1024, stride)[:, : 64 * 64].double().sum(0).view(64, 64)
    return out, part


for layout in ('replicated', 'batched'):
    Bs = 64
    N = Bs * n
    ei = topo.edge_index.to(dev)
    if layout == 'replicated':
        lay = sb.GraphLayout.replicated(ei, n, Bs)
    else:
        lay = sb.GraphLayout.from_batch(sb.graphs.batch_edge_index(topo, Bs).to(dev), (torch.arange(Bs + 1) * n).to(dev), n)
    gen = torch.Generator().manual_seed(5)
    w2 = (torch.randn(64, 64, generator=gen) / 8).to(dev)
    du = torch.randn(N, 64, generator=gen).to(dev)
    mu = torch.randn(N, 64, generator=gen).clamp(min=0).to(dev)
    # fp64 reference: g_j = sum_{i : j in out(i)} du_i  (edge (i, j): row i gathers j)
    bei = sb.graphs.batch_edge_index(topo, Bs).to(dev)
    g64 = torch.zeros(N, 64, dtype=torch.float64, device=dev).index_add_(0, bei[1], du.double()[bei[0]])
    ref_du = (g64 @ w2.double()) * (mu > 0)
    ref_th = g64.t() @ mu.double()
    for flag in (1, 3, 4):
        out, part = one(lay, N, flag, w2, du, mu)
        e1 = float((out.double() - ref_du).abs().max() / ref_du.abs().max())
        e2 = float((part - ref_th).abs().max() / ref_th.abs().max())
        zp = bool(torch.equal(out == 0, ref_du == 0))
        print('%-10s flag %d  du_prev err %.2e  dtheta2 err %.2e  zero pattern %s' % (layout, flag, e1, e2, zp), flush=True)

# ragged tail + accumulate flag are covered by the parity suite (S2V_B200_PATH=tc16); timing
N = B * n
lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B)
w2 = (torch.randn(64, 64) / 8).to(dev); du = torch.randn(N, 64, device=dev); mu = torch.rand(N, 64, device=dev)
out = torch.empty(N, 64, device=dev)
ws = torch.empty(L.s2v_embed_backward_workspace_bytes(7, 64), dtype=torch.uint8, device=dev)
g = lay.c_struct()
for rep in range(2):
    for flag in (3, 4):
        def run(): _lib.check(L.s2v_embed_round_backward(C.byref(g), 7, 64, flag, _p(w2), _p(du), _p(mu), _p(out), _p(ws), ws.numel(), _stream()), 'b')
        for _ in range(3): run()
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
        for _ in range(20): run()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        print('flag %d  %.3f ms  (%.0f GB/s algorithmic, %.2f of 6546)' % (flag, ms, 1276.0 * N / ms / 1e6, 1276.0 * N / ms / 1e6 / 6546.2), flush=True)
