import torch
dev = torch.device('cuda')
x = torch.rand(256 * 1024 * 1024, device=dev); y = torch.empty_like(x)   # 1 GiB each
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
gb = x.numel() * 4 / 1e9
print('read  (sum)   %.0f GB/s' % (gb / timeit(lambda: x.sum()) * 1e3))
print('read  (max)   %.0f GB/s' % (gb / timeit(lambda: x.max()) * 1e3))
print('write (fill)  %.0f GB/s' % (gb / timeit(lambda: y.fill_(1.0)) * 1e3))
print('copy          %.0f GB/s (r+w)' % (2 * gb / timeit(lambda: y.copy_(x)) * 1e3))
print('add 2r+1w     %.0f GB/s' % (3 * gb / timeit(lambda: torch.add(x, y, out=y)) * 1e3))
print('relu r+w      %.0f GB/s' % (2 * gb / timeit(lambda: torch.relu_(y)) * 1e3))
