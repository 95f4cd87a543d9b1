import torch
x = torch.empty(1 << 30, device='cuda', dtype=torch.bfloat16)   # 2 GB
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
ms = t(lambda: x.zero_()); print('memset 2 GB: %.3f ms  %.0f GB/s (write only)' % (ms, 2.147 / ms * 1e3))
ms = t(lambda: y.copy_(x)); print('copy 2 GB:   %.3f ms  %.0f GB/s (read + write)' % (ms, 2 * 2.147 / ms * 1e3))
ms = t(lambda: x.sum()); print('sum 2 GB:    %.3f ms  %.0f GB/s (read only)' % (ms, 2.147 / ms * 1e3))
