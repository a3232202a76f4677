"""Time the tensor-core lin kernels of the GATv2 branch alone (n rows): forward [n,64]x[64,128], dx (+mask), dW."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simmodel_b200 import _lib
from simmodel_b200.layout import _p, _stream
L = _lib.lib()
dev = torch.device("cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 998400
x, w, dy = torch.randn(n, 64, device=dev), torch.randn(128, 64, device=dev) * 0.1, torch.randn(n, 128, device=dev)
y, dx, dw = torch.empty(n, 128, device=dev), torch.empty(n, 64, device=dev), torch.empty(128, 64, device=dev)
ws = torch.empty(L.s2v_rows_reduce64_workspace_bytes(2), dtype=torch.uint8, device=dev)
fns = {"fwd  k_rows_gemm16<1,2,0>": lambda: L.s2v_rows_gemm64(_p(x), 64, 1, _p(w), 2, 0, None, 0, _p(y), 128, n, _stream()),
       "dx   k_rows_gemm16<2,1,0>": lambda: L.s2v_rows_gemm64(_p(dy), 128, 2, _p(w), 1, 1, None, 0, _p(dx), 64, n, _stream()),
       "dx*m k_rows_gemm16<2,1,1>": lambda: L.s2v_rows_gemm64(_p(dy), 128, 2, _p(w), 1, 1, _p(x), 64, _p(dx), 64, n, _stream()),
       "dW   k_rows_reduce16<2>  ": lambda: L.s2v_rows_reduce64(_p(dy), 128, 2, _p(x), 64, _p(dw), _p(ws), ws.numel(), n, _stream()),
       "cuBLAS fwd x@w.T         ": lambda: torch.mm(x, w.t(), out=y),
       "cuBLAS dx dy@w           ": lambda: torch.mm(dy, w, out=dx)}
for name, fn in fns.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%s %7.1f us" % (name, e0.elapsed_time(e1) * 100))
