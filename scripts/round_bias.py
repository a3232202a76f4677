"""One backward round alone: du_prev = (W^T du_t) theta2 * [mu_prev > 0] on both arithmetic paths against fp64:
per-element error and the error of the COLUMN SUMS (what the bias gradients are)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simmodel_b200 as sb
from simmodel_b200 import _lib
from simmodel_b200.layout import _p, _stream
L = _lib.lib()
dev = torch.device("cuda")
topo = sb.graphs.load_topology("NWB_UTR"); B = 5
n = B * topo.num_nodes
layout = sb.GraphLayout.replicated(topo.edge_index.to(dev), topo.num_nodes, B)
gen = torch.Generator().manual_seed(0)
du = torch.randn(n, 64, generator=gen).to(dev)
mu = torch.randn(n, 64, generator=gen).clamp(min=0).to(dev)
w2 = (torch.randn(64, 64, generator=gen) * 0.2).to(dev)
# fp64 reference: g_j = sum_{i: j in out(i)} du_i  (transposed gather), du_prev = (g theta2) * [mu > 0]
ei = sb.graphs.batch_edge_index(topo, B).to(dev)
g = torch.zeros(n, 64, dtype=torch.float64, device=dev).index_add_(0, ei[1], du.double()[ei[0]])
ref = (g @ w2.double()) * (mu > 0)
ws = torch.empty(L.s2v_embed_backward_workspace_bytes(7, 64), dtype=torch.uint8, device=dev)
print("%-6s %14s %14s %14s" % ("path", "max elem err", "colsum err", "colsum err / sum|.|"))
for name, flag in (("ffma", 1), ("tc", 2), ("tcws", 3)):
    out = torch.empty(n, 64, device=dev)
    _lib.check(L.s2v_embed_round_backward(C.byref(layout.c_struct()), 7, 64, flag, _p(w2), _p(du), _p(mu), _p(out), _p(ws), ws.numel(), _stream()), name)
    d = out.double() - ref
    cs = d.sum(0).abs().max()
    print("%-6s %14.3e %14.3e %14.3e" % (name, float(d.abs().max() / ref.abs().max()), float(cs / ref.sum(0).abs().max()), float(cs / ref.abs().sum(0).max())))
    bias = (d / ref.abs().clamp(min=1e-3))[ref.abs() > 1e-3]
    print("        mean signed rel err of elements %.3e, mean |rel err| %.3e, corr with sign(ref) %.3f" % (
        float(bias.mean()), float(bias.abs().mean()), float((torch.sign(d) * torch.sign(ref))[ref.abs() > 1e-3].mean())))
