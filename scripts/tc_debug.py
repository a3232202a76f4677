import ctypes as C, torch, sys
sys.path.insert(0, '.')
from simmodel_b200 import _lib
L = _lib.lib(); dev = torch.device('cuda')
rows = 64
p = lambda t: C.c_void_p(t.data_ptr())
out = torch.zeros(64, 64, device=dev)
def run(A, B):
    _lib.check(L.s2v_debug_tc_gemm(2, p(A), p(B), p(out), rows, C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'dbg')
    torch.cuda.synchronize(); return out.clone()
# B = identity-ish coding: B[r][k] = 1000*r + k  -> out[n][k] = sum_r A[r][n] * B[r][k]
B2 = (torch.arange(rows, device=dev).view(-1, 1) * 1000.0 + torch.arange(64, device=dev).view(1, -1)).float()
for (r0, c0) in [(0, 0), (0, 1), (0, 4), (0, 5), (1, 0), (3, 2), (4, 0), (5, 3), (9, 40), (63, 63)]:
    A2 = torch.zeros(rows, 64, device=dev); A2[r0, c0] = 1.0
    o = run(A2, B2)
    nz = o.abs().sum(1).nonzero().flatten().tolist()
    desc = []
    for n in nz[:4]:
        vals = o[n]
        rr = (vals / 1000).floor(); kk = vals - rr * 1000
        desc.append((n, sorted(set(rr.int().tolist()))[:3], kk[:8].int().tolist()))
    print('A[%d,%d]=1 expect out row %d = B[%d] ->' % (r0, c0, c0, r0), desc)
