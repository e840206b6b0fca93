"""Micro-benchmark of one dense conv layer through the C ABI (for ncu captures and the per-kernel
roofline): python tools/bench_conv.py [fwd|wgrad] B H W Cin Cout k [iters]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(R, 'efficientdet.pytorch_b200')]
import torch
from models import _ops as ops

mode = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
B, H, W, Cin, Cout, k = [int(v) for v in (sys.argv[2:8] if len(sys.argv) >= 8 else (32, 64, 64, 256, 256, 3))]
iters = int(sys.argv[8]) if len(sys.argv) > 8 else 20
dev = torch.device('cuda:0')
w = torch.nn.Parameter(torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5)
bias = torch.randn(Cout, device=dev)
# rotate through enough distinct activations to exceed the 126 MB L2
nbuf = max(2, int(300e6 / (B * H * W * max(Cin, Cout) * 4)) + 1)
xs = [torch.randn(B, H, W, Cin, device=dev) for _ in range(nbuf)]
dys = [torch.randn(B, H, W, Cout, device=dev) for _ in range(nbuf)]
wf, wd = ops.pack_conv(w)
tf, td = ops.pack_conv_tc(w)
dw = torch.zeros_like(w)
def run(i):
    if mode == 'fwd':
        ops.conv2d(xs[i % nbuf], wf, Cout, k, bias=bias, act=1, w_tc=tf if ops.tc_enabled() else None)
    else:
        ops.conv_wgrad(xs[i % nbuf], dys[i % nbuf], dw, None, k, tc=ops.tc_enabled())
for i in range(3): run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters): run(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
fl = 2.0 * B * H * W * k * k * Cin * Cout
print('%s %s precision=%s: %.3f ms/launch, %.1f TFLOP/s (algorithmic fp32-equivalent)' % (mode, (B, H, W, Cin, Cout, k), ops.PRECISION, ms, fl / ms / 1e9))
