import torch
x = torch.empty(1<<30, dtype=torch.int16, device='cuda')  # 2 GiB
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.zero_()); print('memset 2GiB: %.3f ms -> %.0f GB/s write' % (ms, 2*1.0737e9/ms/1e6))
ms = t(lambda: y.copy_(x)); print('copy 2GiB: %.3f ms -> %.0f GB/s read+write' % (ms, 4*1.0737e9/ms/1e6))
ms = t(lambda: x.sum()); print('read 2GiB: %.3f ms -> %.0f GB/s read' % (ms, 2*1.0737e9/ms/1e6))
