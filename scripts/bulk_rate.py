"""cp.async.bulk issue cost and service rate per SM (s2v_debug_bulk_rate): all 148 SMs stream disjoint regions; with
planes > 1 every warp alternates between `planes` regions one plane stride apart (the access pattern of the last
backward stage: T planes of 3 993 600 x 256 B)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
from simmodel_b200 import _lib
L = C.CDLL(_lib.LIB_PATH)
dev = torch.device('cuda')
PLANE = 3993600 * 64
src = torch.rand(5 * PLANE + (1 << 24), device=dev)
out = torch.zeros(2, dtype=torch.int64, device=dev)
for bytes_, warps, planes, stride in ((256, 8, 1, 0), (2048, 1, 1, 0), (2048, 8, 1, 0), (2048, 8, 5, PLANE), (2048, 8, 5, PLANE + 4096 * 3),
                                      (2048, 2, 5, PLANE), (16384, 1, 1, 0), (16384, 1, 5, PLANE), (16384, 2, 5, PLANE)):
    n = min(60, (1 << 20) // bytes_ - 1)
    for _ in range(2):
        rc = L.s2v_debug_bulk_rate(C.c_void_p(src.data_ptr()), C.c_void_p(out.data_ptr()), n, bytes_, warps, planes, C.c_int64(stride), None)
        assert rc == 0, rc
        torch.cuda.synchronize()
    t_issue, t_all = out.tolist()
    print('%6d B x %2d copies, %d issuing warps/SM, %d planes (stride %d floats): %5.0f cycles per issue, all landed after %6d cycles = %.1f B/clk/SM'
          % (bytes_, n, warps, planes, stride, t_issue / n, t_all, n * bytes_ * warps / t_all), flush=True)
