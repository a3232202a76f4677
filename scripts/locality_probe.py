"""How much of the round kernel's time is memory-system locality?  Same kernel, same row count and
degree, different neighbour patterns: real AMS order, RCM-reordered AMS, ring lattice (perfect locality)."""
import ctypes as C, sys, os, numpy as np, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200 import _lib
from simmodel_b200.layout import _p, _stream
import scipy.sparse as sp
from scipy.sparse.csgraph import reverse_cuthill_mckee
L = _lib.lib(); dev = torch.device('cuda'); B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes; ei = topo.edge_index.numpy()
A = sp.csr_matrix((np.ones(ei.shape[1]), (ei[0], ei[1])), shape=(n, n))
perm = reverse_cuthill_mckee(A, symmetric_mode=True); inv = np.empty(n, int); inv[perm] = np.arange(n)
variants = {'ams_natural': ei, 'ams_rcm': np.stack([inv[ei[0]], inv[ei[1]]]),
            'ams_random': np.stack([np.random.default_rng(0).permutation(n)[ei[0]], np.random.default_rng(0).permutation(n)[ei[1]]])}
i = np.arange(n); ring = np.concatenate([np.stack([i, (i + 1) % n]), np.stack([i, (i - 1) % n]), np.stack([i, (i + 2) % n])], 1)
variants['ring3'] = ring[:, :2850] if ring.shape[1] > 2850 else ring
w2 = (torch.randn(64, 64) / 8).to(dev)
for name, e in variants.items():
    order = np.lexsort((e[1], e[0])); e = torch.from_numpy(e[:, order].astype(np.int64)).to(dev)
    lay = sb.GraphLayout.replicated(e, n, B)
    N = B * n
    base = torch.randn(N, 64, device=dev); mu = torch.rand(N, 64, device=dev); out = torch.empty(N, 64, device=dev)
    du = torch.randn(N, 64, device=dev)
    ws = torch.empty(L.s2v_embed_backward_workspace_bytes(7, 64), dtype=torch.uint8, device=dev)
    g = lay.c_struct()
    for flag, nm in ((1, 'ffma'), (2, 'tc'), (3, 'tcws')):
        def fwd(): _lib.check(L.s2v_embed_round_forward(C.byref(g), 64, flag, _p(w2), _p(base), _p(mu), _p(out), _stream()), 'f')
        def bwd(): _lib.check(L.s2v_embed_round_backward(C.byref(g), 7, 64, flag, _p(w2), _p(du), _p(mu), _p(out), _p(ws), ws.numel(), _stream()), 'b')
        res = []
        for fn in (fwd, bwd):
            for _ in range(3): fn()
            torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): fn()
            b.record(); torch.cuda.synchronize(); res.append(a.elapsed_time(b) / 10)
        print('%-12s %-5s E=%d  fwd %.3f ms  bwd %.3f ms' % (name, nm, e.shape[1], res[0], res[1]))
# plain copy of one plane for reference
x = torch.empty(B * n, 64, device=dev); y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
for _ in range(10): y.copy_(x)
b.record(); torch.cuda.synchronize(); t = a.elapsed_time(b) / 10
print('copy of one plane (%.2f GB r+w): %.3f ms = %.0f GB/s' % (2 * x.numel() * 4 / 1e9, t, 2 * x.numel() * 4 / t / 1e6))
