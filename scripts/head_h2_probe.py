"""Q head: accuracy against fp64 (B = 64) and time at B (default 4096) for the kernel selected by S2V_B200_HEAD
(tc16 | h2[,CTAs per SM]).  usage: head_h2_probe.py [B]"""
import os, sys, torch
sys.path.insert(0, '.')
import simmodel_b200 as sb
from simmodel_b200 import ops
dev = torch.device('cuda'); B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes
tag = os.environ.get('S2V_B200_HEAD', '(default)')
gen = torch.Generator().manual_seed(11)
w5 = (torch.randn(1, 128, generator=gen) / 8).to(dev); b5 = torch.randn(1, generator=gen).to(dev)
w6 = (torch.randn(64, 64, generator=gen) / 8).to(dev); b6 = torch.randn(64, generator=gen).to(dev)
w7 = (torch.randn(64, 64, generator=gen) / 8).to(dev); b7 = torch.randn(64, generator=gen).to(dev)
hw = (w5, b5, w6, b6, w7, b7)
with torch.no_grad():
    Bs = 64; N = Bs * n
    lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, Bs)
    mu = (torch.randn(N, 64, generator=gen).clamp(min=0) * torch.rand(N, 1, generator=gen) * 50).to(dev)
    q = ops.s2v_head(lay, mu, hw, True)
    m64 = mu.double().view(Bs, n, 64)
    G = torch.relu(m64.mean(1) @ w6.double().t() + b6.double())                      # [Bs, 64]
    Lr = torch.relu(m64 @ w7.double().t() + b7.double())                             # [Bs, n, 64]
    ref = (G @ w5.double()[0, :64]).unsqueeze(1) + Lr @ w5.double()[0, 64:] + b5.double()
    err = float((q.double().view(Bs, n) - ref).abs().max() / ref.abs().max())
    print('%-10s q err vs fp64 %.2e' % (tag, err), flush=True)
    N = B * n
    lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B)
    mu = torch.rand(N, 64, device=dev)
    for rep in range(2):
        for _ in range(3): ops.s2v_head(lay, mu, hw, True)
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
        for _ in range(20): ops.s2v_head(lay, mu, hw, True)
        b.record(); torch.cuda.synchronize()
        print('%-10s pool + head forward %.3f ms' % (tag, a.elapsed_time(b) / 20), flush=True)
