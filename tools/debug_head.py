import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(R, 'efficientdet.pytorch_b200'), os.path.join(R, 'oracle')]
import torch, effdet_oracle as O
from models.retinahead import RetinaHead
dev = torch.device('cuda:0')
cfg = O.make_config('efficientdet-d0', 20, 64, 2)
sd = O.init_state_dict(cfg, seed=31)
m = RetinaHead(num_classes=20, in_channels=64)
m.load_state_dict({k[len('bbox_head.'):]: v for k, v in sd.items() if k.startswith('bbox_head.')})
m = m.to(dev)
g = torch.Generator().manual_seed(3)
feats = [torch.randn(2, 64, 16 >> i, 16 >> i, generator=g) for i in range(5)]
feats[4] = torch.randn(2, 64, 1, 1, generator=g)
for seed in (3, 4, 5, 6):
  g = torch.Generator().manual_seed(seed)
  feats = [torch.randn(2, 64, 16 >> i, 16 >> i, generator=g) for i in range(5)]
  feats[4] = torch.randn(2, 64, 1, 1, generator=g)
  for which in ('cls',):
      fr = [f.clone().requires_grad_(True) for f in feats]
      fd = [f.to(dev).requires_grad_(True) for f in feats]
      sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in sd.items()}
      cr, rr = O.head_forward(sdg, fr, cfg)
      cd, rd = m(fd)
      g2 = torch.Generator().manual_seed(9)
      lr, l = 0, 0
      pairs = []
      if which in ('cls', 'both'): pairs += list(zip(cd, cr))
      if which in ('reg', 'both'): pairs += list(zip(rd, rr))
      for a, b in pairs:
          wgt = torch.randn(b.shape, generator=g2)
          lr = lr + (b * wgt).sum(); l = l + (a * wgt.to(dev)).sum()
      for p in m.parameters(): p.grad = None
      lr.backward(); l.backward()
      print(which, 'feat grad rel', ['%.2e' % O.rel_err(a.grad.cpu(), b.grad) for a, b in zip(fd, fr)], 'norms', ['%.2e' % float(b.grad.norm()) for b in fr])
      for name, p in m.named_parameters():
          ref = sdg['bbox_head.' + name].grad
          if ref is None or p.grad is None: continue
          pass
