"""Per-layer timing of the depthwise kernels over the D0 block table (bs=32, 512x512 input)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(R, 'efficientdet.pytorch_b200'), os.path.join(R, 'oracle')]
import torch
import effdet_oracle as O
from models import _native as N
dev = torch.device('cuda:0')
cfg = O.make_config('efficientdet-d0', 80, 64, 2)
B = 32
res = 256
tot = {'fwd': 0, 'bwd_data': 0, 'bwd_weight': 0, 'ideal': 0}
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for i, blk in enumerate(cfg['blocks']):
    k, s, C = blk['k'], blk['s'], blk['cin'] * blk['e']
    H = W = res
    pad = O.same_pad(k, s, 224)
    pl, pr, pt, pb = pad
    Ho = (H + pt + pb - k) // s + 1; Wo = (W + pl + pr - k) // s + 1
    x = torch.randn(B, H, W, C, device=dev); dz = torch.randn(B, Ho, Wo, C, device=dev)
    w = torch.randn(k, k, C, device=dev); sc = torch.ones(C, device=dev); sh = torch.zeros(C, device=dev)
    z = torch.empty(B, Ho, Wo, C, device=dev); y = torch.empty_like(z); dx = torch.empty_like(x); dw = torch.zeros(C, 1, k, k, device=dev)
    t_f = timeit(lambda: N.call('effdet_dwconv_fwd', x, x.data_ptr(), w.data_ptr(), sc.data_ptr(), sh.data_ptr(), z.data_ptr(), y.data_ptr(), B, H, W, C, k, s, pt, pl, Ho, Wo))
    t_d = timeit(lambda: N.call('effdet_dwconv_bwd_data', x, dz.data_ptr(), w.data_ptr(), dx.data_ptr(), B, H, W, C, k, s, pt, pl, Ho, Wo))
    t_w = timeit(lambda: N.call('effdet_dwconv_bwd_weight', x, x.data_ptr(), dz.data_ptr(), dw.data_ptr(), B, H, W, C, k, s, pt, pl, Ho, Wo))
    ideal_w = 4e-6 * (x.numel() + dz.numel()) / 6566.7   # ms
    print('blk %2d k%d s%d C%4d %3dx%3d -> %3d : fwd %.3f  bwd_data %.3f  bwd_weight %.3f ms  (bwd_weight HBM ideal %.3f, frac %.2f)' % (i, k, s, C, H, W, Ho, t_f, t_d, t_w, ideal_w, ideal_w / t_w))
    tot['fwd'] += t_f; tot['bwd_data'] += t_d; tot['bwd_weight'] += t_w; tot['ideal'] += ideal_w
    res = Ho
print(tot)
