"""The reference's own modules (baseline/_ref, installed UNMODIFIED by baseline/install_ref.sh) on the host CPU:
used by `bench.py --impl reference` and by bench.py's in-line `cpu_baseline` leg.  Three run-time shims, none of
which touches the reference's arithmetic:
  1. models.efficientnet.load_pretrained_weights -> no-op          (EfficientDet.__init__ downloads weights; no network)
  2. torch.Tensor.cuda -> identity                                 (models/losses.py hard-codes .cuda(); this is the CPU arm)
  3. BiFPNModule.relu1/relu2 -> relu(x).clone()                    (in-place `/=` on a ReLU output breaks autograd on torch>=1.5)
This file must be imported in a process where `models` has NOT been imported from efficientdet.pytorch_b200/ (same
package name: the product is a drop-in)."""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')


def available():
    return os.path.exists(os.path.join(REF, 'models', 'efficientdet.py'))


def load():
    assert available(), 'baseline/_ref missing: run baseline/install_ref.sh in the authoring container'
    assert 'models' not in sys.modules or sys.modules['models'].__file__.startswith(REF), \
        'a different `models` package is already imported in this process'
    sys.path.insert(0, REF)
    import models.efficientnet as ref_effnet
    ref_effnet.load_pretrained_weights = lambda *a, **k: None
    torch.Tensor.cuda = lambda self, *a, **k: self
    import models.bifpn as ref_bifpn
    from models.efficientdet import EfficientDet

    class _ReluClone(torch.nn.Module):
        def forward(self, x):
            return torch.relu(x).clone()

    def build(network, num_classes, W, D, state_dict, is_training=True, threshold=0.01, iou_threshold=0.5):
        m = EfficientDet(num_classes=num_classes, network=network, D_bifpn=D, W_bifpn=W, is_training=is_training,
                         threshold=threshold, iou_threshold=iou_threshold)
        for mod in m.modules():
            if isinstance(mod, ref_bifpn.BiFPNModule):
                mod.relu1, mod.relu2 = _ReluClone(), _ReluClone()
        m.load_state_dict(state_dict)
        return m
    return build


def train_steps(build, network, K, W, D, sd, images, ann, steps, warmup):
    """reference train.py:100-118 inner loop without the optimizer (forward + loss + backward), train mode."""
    m = build(network, K, W, D, sd, is_training=True)
    m.train()
    m.is_training = True
    m.freeze_bn()

    def step():
        for p in m.parameters():
            p.grad = None
        cl, rl = m([images, ann])
        (cl.mean() + rl.mean()).backward()
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / max(steps, 1)


def infer_steps(build, network, K, W, D, sd, image, steps, warmup, threshold, iou_threshold=0.5):
    """reference eval.py:101 call: model(img[1,3,H,W]) -> scores, labels, boxes (NMS = torchvision.ops.nms)."""
    m = build(network, K, W, D, sd, is_training=False, threshold=threshold, iou_threshold=iou_threshold)
    m.eval()
    n = 0
    with torch.no_grad():
        for _ in range(warmup):
            m(image)
        t0 = time.perf_counter()
        for _ in range(steps):
            out = m(image)
            n = int(out[0].numel())
    return (time.perf_counter() - t0) / max(steps, 1), n
