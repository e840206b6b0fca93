"""Time the optimizer step of the reference loop (train.py:115-118: clip_grad_norm_(0.1) + AdamW.step()) on the
EfficientDet-D0 parameter set: torch's own kernels vs the fused two-launch path (models/fused_optim.py)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'efficientdet.pytorch_b200'))
from models import EfficientDet          # noqa: E402
from models import _native as N          # noqa: E402
from models.fused_optim import FusedClipAdamW  # noqa: E402


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device('cuda:0')
    model = EfficientDet(num_classes=20, network='efficientdet-d0').to(dev)
    params = [p for p in model.parameters()]
    g = torch.Generator(device=dev).manual_seed(0)
    for p in params:
        p.grad = torch.randn(p.shape, device=dev, generator=g) * 1e-2
    n = sum(p.numel() for p in params)

    ref = torch.optim.AdamW(params, lr=1e-4)

    def torch_step():
        torch.nn.utils.clip_grad_norm_(params, 0.1)
        ref.step()

    t_ref = timed(torch_step)
    fused = FusedClipAdamW(params, lr=1e-4, max_norm=0.1)
    N.reset_launch_count()
    t_fused = timed(fused.step)
    launches = N.launch_count() / 35
    algo_bytes = 4.0 * n * (1 + 4 + 4)        # norm pass reads g; update reads p,g,m,v and writes p,g,m,v
    print(json.dumps({'tensors': len(params), 'elements': n, 'torch_clip_adamw_ms': round(t_ref, 4),
                      'fused_ms': round(t_fused, 4), 'speedup': round(t_ref / t_fused, 2),
                      'fused_launches_per_step': launches, 'fused_GBps_algorithmic': round(algo_bytes / t_fused / 1e6, 1)}))


if __name__ == '__main__':
    main()
