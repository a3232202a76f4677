"""Cycles per tcgen05 MMA sequence of the backward round, one CTA, tensor pipe otherwise idle."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
from simmodel_b200 import _lib
L = _lib.lib(); dev = torch.device('cuda')
out = torch.zeros(20, device=dev)
for reps in (1, 8, 64):
    _lib.check(L.s2v_debug_tc_gemm(5, None, None, C.c_void_p(out.data_ptr()), reps, C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'rate')
    torch.cuda.synchronize()
    r = out.tolist()
    print('reps %3d: reduce(16 SS MN-major M=128) %.0f cyc | transform A-in-TMEM (24, M=64) %.0f | transform SS K-major (24, M=64) %.0f | reduce+transform_ts %.0f\n          bf16: transform K-major (24, M=64) %.0f | reduce MN-major (24, M=64) %.0f | both %.0f | 8 MN-major M=128 %.0f' % (reps, *r[:8]))
    print('          cycles per bf16 MMA  K-major  M64 N64/128/256: %.0f %.0f %.0f  M128: %.0f %.0f %.0f | MN-major M64: %.0f %.0f %.0f  M128: %.0f %.0f %.0f' % tuple(r[8:20]))
