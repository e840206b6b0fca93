"""HBM bandwidth by access mix on this GPU (CUDA events, best of 10): pure write (fill_), pure read (sum), copy (read+write).
Explains why write-dominated kernels (the MBConv expand convs) sit lower against the copy peak than read-dominated ones."""
import torch
dev = torch.device('cuda:0')
n = 1 << 28                                   # 1 GiB of fp32
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)
def best(fn, nbytes):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return nbytes / min(t) / 1e6
print('pure write  fill_   %.0f GB/s' % best(lambda: a.fill_(1.0), 4 * n))
print('pure read   sum     %.0f GB/s' % best(lambda: a.sum(), 4 * n))
print('copy        copy_   %.0f GB/s (read + write bytes)' % best(lambda: b.copy_(a), 8 * n))
print('1r:6w       expand-like mul: out[6n/7] = f(in[n/7]) ...')
x = torch.empty(n // 8, device=dev)
y = torch.empty(6 * (n // 8), device=dev).view(6, -1)
print('1 read : 6 write    %.0f GB/s' % best(lambda: torch.mul(x.unsqueeze(0), 2.0, out=y) if False else y.copy_(x.unsqueeze(0).expand(6, -1)), 4 * 7 * (n // 8)))
