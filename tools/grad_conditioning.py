"""How far does forward round-off of the dense convolutions move the parameter gradients of this network?

The CPU oracle (fp32 restatement of the reference, oracle/effdet_oracle.py) is run twice on the SAME weights and
batch: once as is, once with every F.conv2d output multiplied by (1 + eps * N(0,1)) -- a model of a convolution
kernel whose products carry `eps` relative error (bf16x3 split products: ~5e-6 per layer measured by
tests/test_gpu_parity.py::test_conv2d_tensor_core_forward_dgrad_wgrad; fp32 atomics order: ~1e-7).  The script
prints the resulting relative change of the losses and of every parameter gradient, i.e. the amplification factor
that turns kernel-level error into end-to-end gradient error.  tests/test_gpu_parity.py::TOL_GRAD is the measured
kernel error times the amplification printed here (worst tensor), with 2x head-room; profiles/ keeps the output.

    python tools/grad_conditioning.py [eps=5e-6] [size=256] [B=2] [seeds=3]
"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(R, 'oracle')]
import torch                          # noqa: E402
import torch.nn.functional as F       # noqa: E402
import effdet_oracle as O             # noqa: E402

eps = float(sys.argv[1]) if len(sys.argv) > 1 else 5e-6
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
seeds = int(sys.argv[4]) if len(sys.argv) > 4 else 3
torch.set_num_threads(os.cpu_count() or 1)
cfg = O.make_config('efficientdet-d0', num_classes=20, W_bifpn=64, D_bifpn=2)
sd = O.init_state_dict(cfg, seed=0)
images, ann = O.synthetic_batch(B, size=size, num_classes=20, seed=200)


def run(noise_seed):
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in sd.items()}
    real = F.conv2d
    g = torch.Generator().manual_seed(noise_seed or 0)

    def noisy(*a, **kw):
        y = real(*a, **kw)
        if noise_seed is None:
            return y
        return y * (1 + eps * torch.randn(y.shape, generator=g))
    O.F.conv2d = noisy
    try:
        cl, rl = O.train_forward(sdg, images, ann, cfg)
        (cl.mean() + rl.mean()).backward()
    finally:
        O.F.conv2d = real
    return float(cl), float(rl), {k: v.grad for k, v in sdg.items() if getattr(v, 'grad', None) is not None}


c0, r0, g0 = run(None)
worst = {}
for s in range(1, seeds + 1):
    c1, r1, g1 = run(s)
    print('noise seed %d: loss rel change cls %.2e reg %.2e' % (s, abs(c1 - c0) / abs(c0), abs(r1 - r0) / abs(r0)))
    for k in g0:
        if float(g0[k].abs().max()) == 0:
            continue
        e = O.rel_err(g1[k], g0[k])
        worst[k] = max(worst.get(k, 0.0), e)
top = sorted(worst.items(), key=lambda kv: -kv[1])[:8]
print('eps = %.1e per conv output; worst gradient changes over %d noise draws (size %d, B %d):' % (eps, seeds, size, B))
for k, e in top:
    print('  %-55s %.3e   amplification %.0fx' % (k, e, e / eps))
med = sorted(worst.values())[len(worst) // 2]
print('median over %d live gradients: %.3e (amplification %.0fx)' % (len(worst), med, med / eps))
