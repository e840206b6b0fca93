"""CPU oracle for the EfficientDet hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this file.  The product (``efficientdet.pytorch_b200/``) never does: it fails loudly when the
CUDA extension is missing instead of falling back to anything in here.

What this is: a functional (state-dict in, tensors out) fp32 restatement, on torch CPU ops,
of the reference model graph toandaominh1997/EfficientDet.Pytorch @ fbe56e5.  Every function
names the reference file:line it follows.  Backward comes from torch autograd over the same
functional graph (the reference itself relies on autograd for everything but swish).

Parity pin: the reference has no golden vectors of its own, so the pin is
``tests/golden/make_golden.py``: it imports the *real* reference from /root/reference in the
authoring container, runs both on identical weights/inputs, asserts bit-exact agreement of
every intermediate, and commits sampled reference outputs under ``tests/golden/`` which
``tests/test_oracle_golden.py`` re-checks anywhere (no /root/reference needed at run time).
NMS: the arithmetic lives in torchvision (unpinned dependency, 0.26.0 here); ``nms_greedy``
restates the published algorithm and is checked against ``torchvision.ops.nms`` on CPU.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Architecture tables
# --------------------------------------------------------------------------------------

# models/utils.py:171-184  (width, depth, nominal resolution, dropout)
_BACKBONE_COEFFS = {
    'efficientnet-b0': (1.0, 1.0, 224), 'efficientnet-b1': (1.0, 1.1, 240),
    'efficientnet-b2': (1.1, 1.2, 260), 'efficientnet-b3': (1.2, 1.4, 300),
    'efficientnet-b4': (1.4, 1.8, 380), 'efficientnet-b5': (1.6, 2.2, 456),
    'efficientnet-b6': (1.8, 2.6, 528), 'efficientnet-b7': (2.0, 3.1, 600),
}
# models/efficientdet.py:10-19
_DET_TO_BACKBONE = {'efficientdet-d%d' % i: 'efficientnet-b%d' % min(i, 6) for i in range(8)}
# models/utils.py:264-269 -- (kernel, repeats, in, out, expand, stride); note stages 5 and 7
# are stride 2 here (upstream EfficientNet has stride 1 there).
_STAGES = [(3, 1, 32, 16, 1, 1), (3, 2, 16, 24, 6, 2), (5, 2, 24, 40, 6, 2), (3, 3, 40, 80, 6, 2),
           (5, 3, 80, 112, 6, 2), (5, 4, 112, 192, 6, 2), (3, 1, 192, 320, 6, 2)]
_SE_RATIO = 0.25
BN_EPS = 1e-3                 # models/utils.py:273-274
DROP_CONNECT_RATE = 0.2       # models/utils.py:260
BIFPN_EPS = 1e-4              # models/bifpn.py:140


def _round_filters(f, width, divisor=8):
    """models/utils.py:54-67."""
    if not width:
        return f
    f = f * width
    nf = max(divisor, int(f + divisor / 2) // divisor * divisor)
    if nf < 0.9 * f:
        nf += divisor
    return int(nf)


def _round_repeats(r, depth):
    """models/utils.py:70-75."""
    return int(math.ceil(depth * r)) if depth else r


def same_pad(k, s, nominal):
    """Static TF-'SAME' pad computed for the *nominal* image size (models/utils.py:126-149).
    Returns (left, right, top, bottom)."""
    out = math.ceil(nominal / s)
    p = max((out - 1) * s + (k - 1) + 1 - nominal, 0)
    return (p // 2, p - p // 2, p // 2, p - p // 2)


def make_config(network='efficientdet-d0', num_classes=80, W_bifpn=64, D_bifpn=2):
    """Expand a network name into the per-block table (models/efficientnet.py:117-178)."""
    bb = _DET_TO_BACKBONE[network]
    width, depth, nominal = _BACKBONE_COEFFS[bb]
    blocks, stage_last, stage_out = [], [], []
    for (k, r, ci, co, e, s) in _STAGES:
        ci, co, r = _round_filters(ci, width), _round_filters(co, width), _round_repeats(r, depth)
        for j in range(r):
            blocks.append(dict(k=k, s=s if j == 0 else 1, cin=ci if j == 0 else co, cout=co, e=e,
                               sq=max(1, int((ci if j == 0 else co) * _SE_RATIO)),
                               # skip rule, models/efficientnet.py:100 (stride of first block of a
                               # stage is the *list* [s] so `== 1` is False even for s == 1)
                               skip=(j > 0)))
        stage_last.append(len(blocks) - 1)
        stage_out.append(co)
    return dict(network=network, backbone=bb, nominal=nominal, stem=_round_filters(32, width),
                head_ch=_round_filters(1280, width), blocks=blocks, stage_last=stage_last,
                stage_out=stage_out, num_classes=num_classes, W=W_bifpn, D=D_bifpn,
                feat=256, stacked=4, A=9, levels=5)


# --------------------------------------------------------------------------------------
# State dict: names/shapes identical to the reference (SURVEY.md section 5 "Checkpoint")
# --------------------------------------------------------------------------------------

def state_dict_spec(cfg):
    """Ordered (name, shape, kind) list; kind in conv|bias|bn_w|bn_b|bn_rm|bn_rv|bn_n|fuse|fc_w|fc_b."""
    spec = []

    def bn(prefix, c):
        spec.extend([(prefix + '.weight', (c,), 'bn_w'), (prefix + '.bias', (c,), 'bn_b'),
                     (prefix + '.running_mean', (c,), 'bn_rm'), (prefix + '.running_var', (c,), 'bn_rv'),
                     (prefix + '.num_batches_tracked', (), 'bn_n')])

    p = 'backbone.'
    spec.append((p + '_conv_stem.weight', (cfg['stem'], 3, 3, 3), 'conv'))
    bn(p + '_bn0', cfg['stem'])
    for i, b in enumerate(cfg['blocks']):
        q = p + '_blocks.%d.' % i
        mid = b['cin'] * b['e']
        if b['e'] != 1:
            spec.append((q + '_expand_conv.weight', (mid, b['cin'], 1, 1), 'conv'))
            bn(q + '_bn0', mid)
        spec.append((q + '_depthwise_conv.weight', (mid, 1, b['k'], b['k']), 'conv'))
        bn(q + '_bn1', mid)
        spec.append((q + '_se_reduce.weight', (b['sq'], mid, 1, 1), 'conv'))
        spec.append((q + '_se_reduce.bias', (b['sq'],), 'bias'))
        spec.append((q + '_se_expand.weight', (mid, b['sq'], 1, 1), 'conv'))
        spec.append((q + '_se_expand.bias', (mid,), 'bias'))
        spec.append((q + '_project_conv.weight', (b['cout'], mid, 1, 1), 'conv'))
        bn(q + '_bn2', b['cout'])
    spec.append((p + '_conv_head.weight', (cfg['head_ch'], cfg['stage_out'][-1], 1, 1), 'conv'))
    bn(p + '_bn1', cfg['head_ch'])
    spec.append((p + '_fc.weight', (1000, cfg['head_ch']), 'fc_w'))
    spec.append((p + '_fc.bias', (1000,), 'fc_b'))
    W, L = cfg['W'], cfg['levels']
    for i, c in enumerate(cfg['stage_out'][-5:]):
        spec.append(('neck.lateral_convs.%d.conv.weight' % i, (W, c, 1, 1), 'conv'))
        spec.append(('neck.lateral_convs.%d.conv.bias' % i, (W,), 'bias'))
    for d in range(cfg['D']):
        q = 'neck.stack_bifpn_convs.%d.' % d
        spec.append((q + 'w1', (2, L), 'fuse'))
        spec.append((q + 'w2', (3, L - 2), 'fuse'))
        for m in range(2 * (L - 1)):
            spec.append((q + 'bifpn_convs.%d.0.conv.weight' % m, (W, W, 3, 3), 'conv'))
            spec.append((q + 'bifpn_convs.%d.0.conv.bias' % m, (W,), 'bias'))
    Fch = cfg['feat']
    for i in range(cfg['stacked']):
        cin = W if i == 0 else Fch
        for t in ('cls', 'reg'):
            spec.append(('bbox_head.%s_convs.%d.conv.weight' % (t, i), (Fch, cin, 3, 3), 'conv'))
            spec.append(('bbox_head.%s_convs.%d.conv.bias' % (t, i), (Fch,), 'bias'))
    spec.append(('bbox_head.retina_cls.weight', (cfg['A'] * cfg['num_classes'], Fch, 3, 3), 'conv'))
    spec.append(('bbox_head.retina_cls.bias', (cfg['A'] * cfg['num_classes'],), 'bias'))
    spec.append(('bbox_head.retina_reg.weight', (cfg['A'] * 4, Fch, 3, 3), 'conv'))
    spec.append(('bbox_head.retina_reg.bias', (cfg['A'] * 4,), 'bias'))
    # the reference registers params in module order: backbone, neck(lateral, stack), bbox_head
    # with cls_convs before reg_convs; reorder head entries accordingly
    head = [s for s in spec if s[0].startswith('bbox_head.')]
    rest = [s for s in spec if not s[0].startswith('bbox_head.')]
    order = (['bbox_head.cls_convs.%d.conv.%s' % (i, w) for i in range(cfg['stacked']) for w in ('weight', 'bias')]
             + ['bbox_head.reg_convs.%d.conv.%s' % (i, w) for i in range(cfg['stacked']) for w in ('weight', 'bias')]
             + ['bbox_head.retina_cls.weight', 'bbox_head.retina_cls.bias',
                'bbox_head.retina_reg.weight', 'bbox_head.retina_reg.bias'])
    hd = {s[0]: s for s in head}
    return rest + [hd[n] for n in order]


# per-layer-kind gains of the 'wellcond' weight set: chosen so every intermediate of D0..D7
# stays O(1) (measured: stage outputs 0.5-3, neck 0.5-3, logits std ~1.5 around the prior bias)
_WELLCOND_GAIN = {'default': 1.0, '_conv_stem': 1.5, '_expand_conv': 1.6, '_depthwise_conv': 1.6,
                  '_project_conv': 1.1, '_project_skip': 0.45, '_se_reduce': 1.0, '_se_expand': 1.0, 'lateral_convs': 1.4,
                  'bifpn_convs': 1.22, 'cls_convs': 1.45, 'reg_convs': 1.45, 'retina_cls': 1.5,
                  'retina_reg': 0.5}


def init_state_dict(cfg, seed=0, mode='wellcond'):
    """Deterministic weights from a CPU generator.

    mode='asbuilt'  : what EfficientDet.__init__ leaves behind (models/efficientdet.py:47-53):
                      every conv ~ N(0, sqrt(2/(k*k*Cout))), BN gamma=1 beta=0, stats (0,1),
                      conv biases 0 (models/module.py:518-525), fusion weights 0.5.
                      Numerically degenerate.
    mode='wellcond' : fan-in scaled convs, random BN affine + running stats, signed fusion
                      weights (exercises the ReLU), prior-probability bias on retina_cls
                      (models/retinahead.py:100-107) -- activations stay O(1) end to end.
    """
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()

    def randn(shape, std=1.0):
        return torch.randn(shape, generator=g, dtype=torch.float32) * std

    def rand(shape, lo, hi):
        return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    for name, shape, kind in state_dict_spec(cfg):
        if kind == 'bn_n':
            sd[name] = torch.zeros((), dtype=torch.long)
            continue
        if mode == 'asbuilt':
            if kind == 'conv':
                t = randn(shape, math.sqrt(2.0 / (shape[2] * shape[3] * shape[0])))
            elif kind in ('bn_w', 'bn_rv'):
                t = torch.ones(shape)
            elif kind in ('bn_b', 'bn_rm', 'bias'):
                t = torch.zeros(shape)
            elif kind == 'fuse':
                t = torch.full(shape, 0.5)
            elif kind == 'fc_w':
                t = rand(shape, -1.0, 1.0) / math.sqrt(shape[1])
            else:
                t = rand(shape, -1.0, 1.0) / math.sqrt(shape[0])
        else:
            if kind == 'conv':
                fan_in = shape[1] * shape[2] * shape[3]
                gain = _WELLCOND_GAIN['default']
                for key, gval in _WELLCOND_GAIN.items():
                    if key in name:
                        gain = gval
                if '_project_conv' in name:
                    # residual (repeat) blocks get a damped branch so deep stages (D4..D7) do not blow up
                    bi = int(name.split('_blocks.')[1].split('.')[0])
                    if cfg['blocks'][bi]['skip']:
                        gain = _WELLCOND_GAIN['_project_skip']
                t = randn(shape, gain / math.sqrt(fan_in))
            elif kind == 'bn_w':
                t = rand(shape, 0.7, 1.3)
            elif kind == 'bn_b':
                t = randn(shape, 0.2)
            elif kind == 'bn_rm':
                t = randn(shape, 0.2)
            elif kind == 'bn_rv':
                t = rand(shape, 0.6, 1.6)
            elif kind == 'bias':
                t = randn(shape, 0.1)
                if name == 'bbox_head.retina_cls.bias':
                    t = t + float(-np.log((1 - 0.01) / 0.01))
            elif kind == 'fuse':
                t = rand(shape, -0.3, 1.2)
            elif kind == 'fc_w':
                t = randn(shape, 1.0 / math.sqrt(shape[1]))
            else:
                t = torch.zeros(shape)
        sd[name] = t
    return sd


# --------------------------------------------------------------------------------------
# Backbone (models/efficientnet.py, models/utils.py)
# --------------------------------------------------------------------------------------

def swish(x):
    """models/utils.py:31-47 -- x * sigmoid(x)."""
    return x * torch.sigmoid(x)


def _bn(sd, p, x):
    """nn.BatchNorm2d in eval mode (frozen; models/efficientdet.py:88-92), eps 1e-3."""
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'],
                        sd[p + '.bias'], False, 0.0, BN_EPS)


def _same_conv(x, w, bias, k, s, nominal, groups=1):
    """Conv2dStaticSamePadding.forward (models/utils.py:151-155): ZeroPad2d then conv pad 0."""
    pad = same_pad(k, s, nominal)
    if any(pad):
        x = F.pad(x, pad)
    return F.conv2d(x, w, bias, s, 0, 1, groups)


def mbconv_forward(sd, q, blk, x, nominal, keep=None, drop_p=0.0, collect=None):
    """MBConvBlock.forward (models/efficientnet.py:75-105).

    keep : optional [B,1,1,1] tensor of uniform samples in [0,1) -- the torch.rand of
           drop_connect (models/utils.py:79-90); None disables drop-connect (eval mode)."""
    inp = x
    mid = blk['cin'] * blk['e']
    if blk['e'] != 1:
        x = swish(_bn(sd, q + '_bn0', _same_conv(x, sd[q + '_expand_conv.weight'], None, 1, 1, nominal)))
    x = swish(_bn(sd, q + '_bn1', _same_conv(x, sd[q + '_depthwise_conv.weight'], None,
                                              blk['k'], blk['s'], nominal, groups=mid)))
    sq = F.adaptive_avg_pool2d(x, 1)
    sq = _same_conv(sq, sd[q + '_se_reduce.weight'], sd[q + '_se_reduce.bias'], 1, 1, nominal)
    sq = _same_conv(swish(sq), sd[q + '_se_expand.weight'], sd[q + '_se_expand.bias'], 1, 1, nominal)
    x = torch.sigmoid(sq) * x
    x = _bn(sd, q + '_bn2', _same_conv(x, sd[q + '_project_conv.weight'], None, 1, 1, nominal))
    if blk['skip'] and blk['cin'] == blk['cout']:
        if keep is not None and drop_p:
            kp = 1 - drop_p
            x = x / kp * torch.floor(kp + keep)
        x = x + inp
    return x


def backbone_forward(sd, x, cfg, prefix='backbone.', keep_samples=None):
    """EfficientNet.extract_features (models/efficientnet.py:190-209): 7 stage outputs.

    keep_samples: None (eval) or list with one [B,1,1,1] uniform tensor per *skip* block, in
    block order (only blocks with a skip consume a torch.rand; block 0 has rate 0 -> falsy)."""
    n = cfg['nominal']
    x = swish(_bn(sd, prefix + '_bn0', _same_conv(x, sd[prefix + '_conv_stem.weight'], None, 3, 2, n)))
    outs, ki = [], 0
    nb = len(cfg['blocks'])
    for i, blk in enumerate(cfg['blocks']):
        rate = DROP_CONNECT_RATE * float(i) / nb
        keep = None
        if keep_samples is not None and blk['skip'] and blk['cin'] == blk['cout'] and rate:
            keep = keep_samples[ki]
            ki += 1
        x = mbconv_forward(sd, prefix + '_blocks.%d.' % i, blk, x, n, keep=keep, drop_p=rate)
        if i in cfg['stage_last']:
            outs.append(x)
    return outs


# --------------------------------------------------------------------------------------
# Neck (models/bifpn.py)
# --------------------------------------------------------------------------------------

def bifpn_layer_forward(sd, q, feats, eps=BIFPN_EPS, collect=None):
    """BiFPNModule.forward (models/bifpn.py:172-203).  feats: list of L maps, fine -> coarse."""
    L = len(feats)
    w1 = torch.relu(sd[q + 'w1'])
    w1 = w1 / (torch.sum(w1, dim=0) + eps)
    w2 = torch.relu(sd[q + 'w2'])
    w2 = w2 / (torch.sum(w2, dim=0) + eps)

    def conv(idx, t):
        return F.conv2d(t, sd[q + 'bifpn_convs.%d.0.conv.weight' % idx],
                        sd[q + 'bifpn_convs.%d.0.conv.bias' % idx], 1, 1)

    path = list(feats)
    orig = list(feats)
    idx = 0
    for i in range(L - 1, 0, -1):                       # top-down  (:188-192)
        t = (w1[0, i - 1] * path[i - 1] + w1[1, i - 1] * F.interpolate(path[i], scale_factor=2, mode='nearest')) \
            / (w1[0, i - 1] + w1[1, i - 1] + eps)
        path[i - 1] = conv(idx, t)
        idx += 1
    for i in range(0, L - 2):                           # bottom-up (:194-198)
        t = (w2[0, i] * path[i + 1] + w2[1, i] * F.max_pool2d(path[i], kernel_size=2) + w2[2, i] * orig[i + 1]) \
            / (w2[0, i] + w2[1, i] + w2[2, i] + eps)
        path[i + 1] = conv(idx, t)
        idx += 1
    t = (w1[0, L - 1] * path[L - 1] + w1[1, L - 1] * F.max_pool2d(path[L - 2], kernel_size=2)) \
        / (w1[0, L - 1] + w1[1, L - 1] + eps)           # top level (:200-202)
    path[L - 1] = conv(idx, t)
    return path


def bifpn_forward(sd, feats, cfg, prefix='neck.', collect=None):
    """BIFPN.forward (models/bifpn.py:96-129): 5 lateral 1x1 convs then D stacked layers."""
    lat = [F.conv2d(f, sd[prefix + 'lateral_convs.%d.conv.weight' % i],
                    sd[prefix + 'lateral_convs.%d.conv.bias' % i]) for i, f in enumerate(feats)]
    if collect is not None:
        collect['laterals'] = list(lat)
    for d in range(cfg['D']):
        lat = bifpn_layer_forward(sd, prefix + 'stack_bifpn_convs.%d.' % d, lat)
        if collect is not None:
            collect['bifpn%d' % d] = list(lat)
    return lat


# --------------------------------------------------------------------------------------
# Head (models/retinahead.py)
# --------------------------------------------------------------------------------------

def head_forward_single(sd, x, cfg, prefix='bbox_head.'):
    """RetinaHead.forward_single (models/retinahead.py:109-129)."""
    c, r = x, x
    for i in range(cfg['stacked']):
        c = torch.relu(F.conv2d(c, sd[prefix + 'cls_convs.%d.conv.weight' % i],
                                sd[prefix + 'cls_convs.%d.conv.bias' % i], 1, 1))
    for i in range(cfg['stacked']):
        r = torch.relu(F.conv2d(r, sd[prefix + 'reg_convs.%d.conv.weight' % i],
                                sd[prefix + 'reg_convs.%d.conv.bias' % i], 1, 1))
    cls = torch.sigmoid(F.conv2d(c, sd[prefix + 'retina_cls.weight'], sd[prefix + 'retina_cls.bias'], 1, 1))
    B = x.shape[0]
    cls = cls.permute(0, 2, 3, 1).contiguous().view(B, -1, cfg['num_classes'])
    reg = F.conv2d(r, sd[prefix + 'retina_reg.weight'], sd[prefix + 'retina_reg.bias'], 1, 1)
    reg = reg.permute(0, 2, 3, 1).contiguous().view(B, -1, 4)
    return cls, reg


def head_forward(sd, feats, cfg, prefix='bbox_head.'):
    """RetinaHead.forward (models/retinahead.py:131-132): same weights on every level."""
    outs = [head_forward_single(sd, f, cfg, prefix) for f in feats]
    return [o[0] for o in outs], [o[1] for o in outs]


# --------------------------------------------------------------------------------------
# Anchors / box coding (models/module.py)
# --------------------------------------------------------------------------------------

def anchors_for(height, width, levels=(3, 4, 5, 6, 7)):
    """Anchors.forward + generate_anchors + shift (models/module.py:161-214,252-273).
    float64 NumPy arithmetic, one final cast to fp32; returns np.float32 [1, A, 4]."""
    ratios = np.array([0.5, 1, 2])
    scales = np.array([2 ** 0, 2 ** (1.0 / 3.0), 2 ** (2.0 / 3.0)])
    rows = []
    for lv in levels:
        stride, base = 2 ** lv, 2 ** (lv + 2)
        fh, fw = (height + stride - 1) // stride, (width + stride - 1) // stride
        # 9 base boxes: ratio-major, scale-minor
        wh = base * np.tile(scales, (2, len(ratios))).T
        areas = wh[:, 0] * wh[:, 1]
        w = np.sqrt(areas / np.repeat(ratios, len(scales)))
        h = w * np.repeat(ratios, len(scales))
        base_boxes = np.zeros((9, 4))
        base_boxes[:, 2] = w
        base_boxes[:, 3] = h
        base_boxes[:, 0::2] -= np.tile(w * 0.5, (2, 1)).T
        base_boxes[:, 1::2] -= np.tile(h * 0.5, (2, 1)).T
        sx = (np.arange(0, fw) + 0.5) * stride
        sy = (np.arange(0, fh) + 0.5) * stride
        sx, sy = np.meshgrid(sx, sy)
        shifts = np.vstack((sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel())).transpose()
        boxes = (base_boxes.reshape((1, 9, 4)) + shifts.reshape((1, -1, 4)).transpose((1, 0, 2))).reshape((-1, 4))
        rows.append(boxes)
    allb = np.zeros((0, 4)).astype(np.float32)
    for r in rows:
        allb = np.append(allb, r, axis=0)
    return np.expand_dims(allb, axis=0).astype(np.float32)


def decode_boxes(anchors, deltas):
    """BBoxTransform.forward (models/module.py:24-49), std (0.1,0.1,0.2,0.2), mean 0."""
    std = torch.from_numpy(np.array([0.1, 0.1, 0.2, 0.2]).astype(np.float32))
    mean = torch.from_numpy(np.array([0, 0, 0, 0]).astype(np.float32))
    w = anchors[:, :, 2] - anchors[:, :, 0]
    h = anchors[:, :, 3] - anchors[:, :, 1]
    cx = anchors[:, :, 0] + 0.5 * w
    cy = anchors[:, :, 1] + 0.5 * h
    dx = deltas[:, :, 0] * std[0] + mean[0]
    dy = deltas[:, :, 1] * std[1] + mean[1]
    dw = deltas[:, :, 2] * std[2] + mean[2]
    dh = deltas[:, :, 3] * std[3] + mean[3]
    pcx, pcy = cx + dx * w, cy + dy * h
    pw, ph = torch.exp(dw) * w, torch.exp(dh) * h
    return torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], dim=2)


def clip_boxes(boxes, height, width):
    """ClipBoxes.forward (models/module.py:57-67)."""
    b = boxes.clone()
    b[:, :, 0] = torch.clamp(b[:, :, 0], min=0)
    b[:, :, 1] = torch.clamp(b[:, :, 1], min=0)
    b[:, :, 2] = torch.clamp(b[:, :, 2], max=width)
    b[:, :, 3] = torch.clamp(b[:, :, 3], max=height)
    return b


# --------------------------------------------------------------------------------------
# Loss (models/losses.py)
# --------------------------------------------------------------------------------------

def pairwise_iou(a, b):
    """calc_iou (models/losses.py:6-26): [A,4] x [G,4] -> [A,G]; union clamped >= 1e-8."""
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    iw = torch.min(a[:, 2].unsqueeze(1), b[:, 2]) - torch.max(a[:, 0].unsqueeze(1), b[:, 0])
    ih = torch.min(a[:, 3].unsqueeze(1), b[:, 3]) - torch.max(a[:, 1].unsqueeze(1), b[:, 1])
    iw = torch.clamp(iw, min=0)
    ih = torch.clamp(ih, min=0)
    ua = ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])).unsqueeze(1) + area_b - iw * ih
    ua = torch.clamp(ua, min=1e-8)
    return (iw * ih) / ua


def focal_loss(cls, reg, anchors, annots, alpha=0.25, gamma=2.0):
    """FocalLoss.forward (models/losses.py:32-152).  Returns (cls_loss[1], reg_loss[1])."""
    B = cls.shape[0]
    a = anchors[0]
    aw, ah = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
    acx, acy = a[:, 0] + 0.5 * aw, a[:, 1] + 0.5 * ah
    cls_losses, reg_losses = [], []
    for j in range(B):
        gt = annots[j]
        gt = gt[gt[:, 4] != -1]
        if gt.shape[0] == 0:                                       # :54-58
            reg_losses.append(torch.tensor(0).float())
            cls_losses.append(torch.tensor(0).float())
            continue
        p = torch.clamp(cls[j], 1e-4, 1.0 - 1e-4)                  # :60
        iou = pairwise_iou(a, gt[:, :4])
        iou_max, iou_arg = torch.max(iou, dim=1)                   # :65
        tgt = torch.ones(p.shape) * -1
        tgt[torch.lt(iou_max, 0.4), :] = 0                         # :74
        pos = torch.ge(iou_max, 0.5)
        npos = pos.sum()
        assigned = gt[iou_arg, :]
        tgt[pos, :] = 0
        tgt[pos, assigned[pos, 4].long()] = 1
        af = torch.ones(tgt.shape) * alpha
        af = torch.where(torch.eq(tgt, 1.), af, 1. - af)
        fw = torch.where(torch.eq(tgt, 1.), 1. - p, p)
        fw = af * torch.pow(fw, gamma)
        bce = -(tgt * torch.log(p) + (1.0 - tgt) * torch.log(1.0 - p))
        l = fw * bce
        l = torch.where(torch.ne(tgt, -1.0), l, torch.zeros(l.shape))
        cls_losses.append(l.sum() / torch.clamp(npos.float(), min=1.0))
        if pos.sum() > 0:                                          # :108-148
            asg = assigned[pos, :]
            gw, gh = asg[:, 2] - asg[:, 0], asg[:, 3] - asg[:, 1]
            gcx, gcy = asg[:, 0] + 0.5 * gw, asg[:, 1] + 0.5 * gh
            gw, gh = torch.clamp(gw, min=1), torch.clamp(gh, min=1)
            t = torch.stack(((gcx - acx[pos]) / aw[pos], (gcy - acy[pos]) / ah[pos],
                             torch.log(gw / aw[pos]), torch.log(gh / ah[pos]))).t()
            t = t / torch.Tensor([[0.1, 0.1, 0.2, 0.2]])
            d = torch.abs(t - reg[j][pos, :])
            rl = torch.where(torch.le(d, 1.0 / 9.0), 0.5 * 9.0 * torch.pow(d, 2), d - 0.5 / 9.0)
            reg_losses.append(rl.mean())
        else:
            reg_losses.append(torch.tensor(0).float())
    return (torch.stack(cls_losses).mean(dim=0, keepdim=True),
            torch.stack(reg_losses).mean(dim=0, keepdim=True))


# --------------------------------------------------------------------------------------
# NMS (torchvision.ops.nms semantics; call site models/efficientdet.py:82-83)
# --------------------------------------------------------------------------------------

def nms_greedy(boxes, scores, thr):
    """Greedy class-agnostic NMS restating torchvision's CPU kernel: stable descending sort
    (ties -> lower index first), area = (x2-x1)*(y2-y1), suppress iff IoU > thr (strict),
    all arithmetic in fp32.  Returns int64 indices in descending-score order."""
    b = boxes.detach().cpu().numpy().astype(np.float32)
    s = scores.detach().cpu().numpy().astype(np.float32)
    n = b.shape[0]
    order = np.argsort(-s, kind='stable')
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    dead = np.zeros(n, dtype=bool)
    keep = []
    thr = np.float32(thr)
    for ii in range(n):
        i = order[ii]
        if dead[i]:
            continue
        keep.append(i)
        rest = order[ii + 1:]
        xx1 = np.maximum(x1[i], x1[rest]); yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest]); yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), xx2 - xx1); h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide='ignore', invalid='ignore'):
            ovr = inter / (areas[i] + areas[rest] - inter)
        dead[rest[ovr > thr]] = True
    return torch.from_numpy(np.asarray(keep, dtype=np.int64))


# --------------------------------------------------------------------------------------
# Detector wrapper (models/efficientdet.py)
# --------------------------------------------------------------------------------------

def features(sd, images, cfg, keep_samples=None, collect=None):
    """EfficientDet.extract_feat (models/efficientdet.py:94-100)."""
    P = backbone_forward(sd, images, cfg, keep_samples=keep_samples)
    if collect is not None:
        collect['P'] = list(P)
    return bifpn_forward(sd, P[-5:], cfg, collect=collect)


def raw_outputs(sd, images, cfg, keep_samples=None, collect=None):
    """backbone -> neck -> head -> cat over levels (models/efficientdet.py:62-66)."""
    feats = features(sd, images, cfg, keep_samples, collect)
    cls_l, reg_l = head_forward(sd, feats, cfg)
    cls, reg = torch.cat(cls_l, dim=1), torch.cat(reg_l, dim=1)
    anchors = torch.from_numpy(anchors_for(images.shape[2], images.shape[3]))
    if collect is not None:
        collect.update(neck=list(feats), cls=cls, reg=reg, anchors=anchors)
    return cls, reg, anchors


def train_forward(sd, images, annots, cfg, keep_samples=None, collect=None):
    """EfficientDet.forward, is_training=True (models/efficientdet.py:57-68)."""
    cls, reg, anchors = raw_outputs(sd, images, cfg, keep_samples, collect)
    return focal_loss(cls, reg, anchors, annots)


def detect(sd, image, cfg, threshold=0.01, iou_threshold=0.5, nms_fn=None, collect=None):
    """EfficientDet.forward, is_training=False (models/efficientdet.py:69-86); image 0 only."""
    cls, reg, anchors = raw_outputs(sd, image, cfg, None, collect)
    boxes = clip_boxes(decode_boxes(anchors, reg), image.shape[2], image.shape[3])
    scores = torch.max(cls, dim=2, keepdim=True)[0]
    mask = (scores > threshold)[0, :, 0]
    if mask.sum() == 0:
        return [torch.zeros(0), torch.zeros(0), torch.zeros(0, 4)]
    cls_f, boxes_f, scores_f = cls[:, mask, :], boxes[:, mask, :], scores[:, mask, :]
    keep = (nms_fn or nms_greedy)(boxes_f[0], scores_f[0, :, 0], iou_threshold)
    sc, cl = cls_f[0, keep, :].max(dim=1)
    if collect is not None:
        collect.update(boxes=boxes, mask=mask, keep=keep)
    return [sc, cl, boxes_f[0, keep, :]]


# --------------------------------------------------------------------------------------
# Synthetic workload
# --------------------------------------------------------------------------------------

def synthetic_batch(B, size=512, G=8, num_classes=80, seed=0, empty_first=False):
    """images ~ randn (seed), annotations in `collater` format [B,G,5], pad rows = -1."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(B, 3, size, size, generator=g, dtype=torch.float32)
    g2 = torch.Generator().manual_seed(seed + 1)
    ann = torch.full((B, G, 5), -1.0)
    for b in range(B):
        n = int(torch.randint(1, G + 1, (1,), generator=g2))
        if empty_first and b == 0:
            n = 0
        for i in range(n):
            x1 = float(torch.rand(1, generator=g2)) * size * 0.75
            y1 = float(torch.rand(1, generator=g2)) * size * 0.75
            w = 8 + float(torch.rand(1, generator=g2)) * (size * 0.25 - 8)
            h = 8 + float(torch.rand(1, generator=g2)) * (size * 0.25 - 8)
            lab = int(torch.randint(0, num_classes, (1,), generator=g2))
            ann[b, i] = torch.tensor([x1, y1, x1 + w, y1 + h, float(lab)])
    return images, ann


def rel_err(a, b):
    """||a-b||_2 / ||b||_2 in float64 (per-tensor norm-relative metric)."""
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    d = torch.linalg.vector_norm(a - b)
    n = torch.linalg.vector_norm(b)
    return float(d / n) if n > 0 else float(d)
