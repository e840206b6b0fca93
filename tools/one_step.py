"""One (or N) EfficientDet-D0 512x512 bs=32 training steps in train mode, for profiling under ncu.
usage: python tools/one_step.py [steps=1] [network=efficientdet-d0] [size=512] [bs=32] [W=64] [D=2]"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(R, 'oracle'), os.path.join(R, 'efficientdet.pytorch_b200')]
import torch                        # noqa: E402
import effdet_oracle as O           # noqa: E402
from models import EfficientDet     # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
net = sys.argv[2] if len(sys.argv) > 2 else 'efficientdet-d0'
size = int(sys.argv[3]) if len(sys.argv) > 3 else 512
bs = int(sys.argv[4]) if len(sys.argv) > 4 else 32
W = int(sys.argv[5]) if len(sys.argv) > 5 else 64
D = int(sys.argv[6]) if len(sys.argv) > 6 else 2
dev = torch.device('cuda:0')
cfg = O.make_config(net, 80, W, D)
m = EfficientDet(num_classes=80, network=net, D_bifpn=D, W_bifpn=W, is_training=True)
m.load_state_dict(O.init_state_dict(cfg, seed=0))
m = m.to(dev)
m.train()
m.is_training = True
m.freeze_bn()
images, ann = O.synthetic_batch(bs, size=size, G=8, num_classes=80, seed=1000)
images, ann = images.to(dev), ann.to(dev)
for _ in range(steps):
    for p in m.parameters():
        p.grad = None
    cl, rl = m([images, ann])
    (cl.mean() + rl.mean()).backward()
torch.cuda.synchronize()
print('done', float(cl), float(rl))
