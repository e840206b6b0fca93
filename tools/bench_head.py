"""Micro-benchmark of ONE RetinaHead tower layer over the five pyramid levels (the multi-level launches of
models/_ops.py::RetinaHeadFn) through the C ABI: forward (bias + ReLU), data gradient (ReLU mask) and weight
gradient.  Staged kernel variants are selected by the environment before the library is loaded, e.g.

    python tools/bench_head.py                      # validated kernels
    EFFDET_B200_COAL=1 python tools/bench_head.py   # coalesced epilogue
    EFFDET_B200_PAIR=1 python tools/bench_head.py   # CTA-pair kernels

Every run also prints checksums of the three results so that runs under different flags can be compared.
usage: python tools/bench_head.py [B] [Cin] [Cout] [size] [iters]"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(R, 'efficientdet.pytorch_b200')]
import torch                      # noqa: E402
from models import _native as N   # noqa: E402
from models import _ops as ops    # noqa: E402

B, Cin, Cout, size, iters = [int(v) for v in (sys.argv[1:6] + ['32', '256', '256', '512', '10'][len(sys.argv) - 1:])]
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(0)
sides = [size >> s for s in (3, 4, 5, 6, 7)]
w = torch.nn.Parameter(torch.randn(Cout, Cin, 3, 3, device=dev, generator=g) / (Cin * 9) ** 0.5)
bias = torch.randn(Cout, device=dev, generator=g)
nbuf = 3                                                        # 3 x ~0.7 GB of activations: every launch misses L2
xs = [[torch.randn(B, s, s, Cin, device=dev, generator=g) for s in sides] for _ in range(nbuf)]
dys = [[torch.randn(B, s, s, Cout, device=dev, generator=g) for s in sides] for _ in range(nbuf)]
wf, wd = ops.pack_conv(w)
tf, td = ops.pack_conv_tc(w)
tc = ops.tc_enabled()
dw = torch.zeros_like(w)
db = torch.zeros(Cout, device=dev)


def fwd(i):
    return ops.conv2d_multi(xs[i % nbuf], wf, Cout, 3, bias=bias, act=N.ACT_RELU, w_tc=tf if tc else None)


def dgrad(i):
    return ops.conv2d_multi(dys[i % nbuf], wd, Cin, 3, w_tc=td if tc else None, masks=xs[i % nbuf])


def wgrad(i):
    lv = [dict(x_ptr=N.f32(x), x_bs=x.shape[1] * x.shape[2] * Cin, dy_ptr=N.f32(d), dy_bs=d.shape[1] * d.shape[2] * Cout,
               B=B, H=x.shape[1], W=x.shape[2]) for x, d in zip(xs[i % nbuf], dys[i % nbuf])]
    ops.conv_wgrad_multi(xs[0][0], lv, dw, db, Cin, Cout, 3, tc=tc)
    return [dw]


def timed(fn):
    for i in range(2):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


px = B * sum(s * s for s in sides)
flops = 2.0 * px * 9 * Cin * Cout
flags = {k: os.environ.get(k, '0') for k in ('EFFDET_B200_COAL', 'EFFDET_B200_PAIR')}
print('levels', sides, 'B', B, '%d->%d' % (Cin, Cout), 'precision', ops.PRECISION, flags)
for name, fn in (('fwd', fwd), ('dgrad', dgrad), ('wgrad', wgrad)):
    if name == 'wgrad':
        dw.zero_(); db.zero_()
        out = fn(0)
        chk = [float(dw.double().sum()), float(dw.double().abs().sum()), float(db.double().sum())]
    else:
        out = fn(0)
        chk = [float(sum(o.double().sum() for o in out)), float(sum(o.double().abs().sum() for o in out))]
    ms = timed(fn)
    print('%-6s %.3f ms  %.1f TFLOP/s algorithmic   checksum %s' % (name, ms, flops / ms / 1e9, ['%.6e' % c for c in chk]))
