"""Per-kernel device time of the benchmarked step (AMS, B graphs, DQN loss) from a CUPTI trace (torch.profiler), for one
or more library builds: kernel_times.py B lib1.so [lib2.so ...]   (each library in its own process)."""
import os, re, subprocess, sys
if len(sys.argv) > 2 and sys.argv[1] != '--child':
    for lib in sys.argv[2:]:
        subprocess.run([sys.executable, __file__, '--child', sys.argv[1], lib], check=True)
    sys.exit(0)
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, '.')
B, lib = int(sys.argv[2]), sys.argv[3]
from simmodel_b200 import _lib
_lib.LIB_PATH = os.path.abspath(lib)
import simmodel_b200 as sb
dev = torch.device('cuda')
topo = sb.graphs.load_topology('NWB_AMS'); n = topo.num_nodes
lay = sb.GraphLayout.replicated(topo.edge_index.to(dev), n, B); N = B * n
torch.manual_seed(0)
net = sb.QNet({"emb_dim": 64, "emb_iter_T": 5, "norm_agg": True, "node_dim": 7}, verbose=False).to(dev)
x = torch.rand(N, 7, device=dev)
sel = torch.arange(B, device=dev) * n + torch.randint(0, n, (B,), device=dev)
tgt = torch.randn(B, device=dev)
def step():
    net.zero_grad(set_to_none=True)
    q = net.forward_graph(x, lay)
    torch.nn.functional.smooth_l1_loss(q[sel], tgt).backward()
for _ in range(3): step()
torch.cuda.synchronize()
R = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(R): step()
    torch.cuda.synchronize()
rows = {}
for e in prof.events():
    m = re.search(r'\bk_\w+(<[^>(]*>)?', e.name)
    if 'cuda' in str(getattr(e, 'device_type', '')).lower() and m:
        nm = m.group(0)
        t = getattr(e, 'device_time', None) or getattr(e, 'cuda_time', 0)
        a = rows.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in rows.values())
print('== %s  B=%d: %.3f ms of own kernels per step' % (lib, B, tot / R / 1e3))
for nm, (c, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:12]:
    print('  %-58s x%-3d %8.1f us each  %5.1f %%' % (nm, c // R, t / c, 100 * t / tot))
