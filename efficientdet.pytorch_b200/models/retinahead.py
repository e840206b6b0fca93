"""RetinaNet-style head of the reference (``models/retinahead.py``), B200-native.

Two towers of ``stacked_convs`` x (3x3 conv 256 + bias + ReLU), then ``retina_cls`` (3x3 -> A*K,
sigmoid) and ``retina_reg`` (3x3 -> A*4); the same weights run on every pyramid level
(reference :109-132).  All levels run inside one autograd node; the last convs write straight into
the concatenated ``[B, sum(HWA), K]`` / ``[B, sum(HWA), 4]`` buffers (sigmoid fused in the epilogue),
so the reference's permute / contiguous / view / torch.cat chain (:117-128, efficientdet.py:64-65)
costs nothing.
"""
import numpy as np
import torch.nn as nn

from . import _ops
from .module import ConvModule, bias_init_with_prob, normal_init


def multi_apply(func, *args, **kwargs):
    from functools import partial
    pfunc = partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


class RetinaHead(nn.Module):
    def __init__(self, num_classes, in_channels, feat_channels=256, anchor_scales=[8, 16, 32],
                 anchor_ratios=[0.5, 1.0, 2.0], anchor_strides=[4, 8, 16, 32, 64], stacked_convs=4,
                 octave_base_scale=4, scales_per_octave=3, conv_cfg=None, norm_cfg=None, **kwargs):
        super().__init__()
        self.in_channels, self.num_classes, self.feat_channels = in_channels, num_classes, feat_channels
        self.anchor_scales, self.anchor_ratios, self.anchor_strides = anchor_scales, anchor_ratios, anchor_strides
        self.stacked_convs = stacked_convs
        self.octave_base_scale, self.scales_per_octave = octave_base_scale, scales_per_octave
        self.conv_cfg, self.norm_cfg = conv_cfg, norm_cfg
        self.cls_out_channels = num_classes
        self.num_anchors = len(self.anchor_ratios) * len(self.anchor_scales)
        self._init_layers()

    def _init_layers(self):
        self.relu = nn.ReLU(inplace=True)
        self.cls_convs = nn.ModuleList()
        self.reg_convs = nn.ModuleList()
        for i in range(self.stacked_convs):
            chn = self.in_channels if i == 0 else self.feat_channels
            self.cls_convs.append(ConvModule(chn, self.feat_channels, 3, stride=1, padding=1,
                                             conv_cfg=self.conv_cfg, norm_cfg=self.norm_cfg))
            self.reg_convs.append(ConvModule(chn, self.feat_channels, 3, stride=1, padding=1,
                                             conv_cfg=self.conv_cfg, norm_cfg=self.norm_cfg))
        self.retina_cls = nn.Conv2d(self.feat_channels, self.num_anchors * self.cls_out_channels, 3, padding=1)
        self.retina_reg = nn.Conv2d(self.feat_channels, self.num_anchors * 4, 3, padding=1)
        self.output_act = nn.Sigmoid()

    def init_weights(self):
        for m in self.cls_convs:
            normal_init(m.conv, std=0.01)
        for m in self.reg_convs:
            normal_init(m.conv, std=0.01)
        normal_init(self.retina_cls, std=0.01, bias=bias_init_with_prob(0.01))
        normal_init(self.retina_reg, std=0.01)

    def _params(self):
        P = []
        for m in self.cls_convs:
            P += [m.conv.weight, m.conv.bias]
        for m in self.reg_convs:
            P += [m.conv.weight, m.conv.bias]
        return P + [self.retina_cls.weight, self.retina_cls.bias, self.retina_reg.weight, self.retina_reg.bias]

    def forward_concat_nhwc(self, feats):
        """NHWC feature maps -> (cls [B, sum(HWA), K] probabilities, reg [B, sum(HWA), 4])."""
        params = self._params()
        K, A = self.num_classes, self.num_anchors
        Kp = (K + 3) // 4 * 4
        if Kp != K:
            # the kernels move channels in vectors of 4: a class count that is not a multiple of 4 (custom datasets with
            # 1, 3, 90 ... classes; the reference accepts any) runs with zero-weight dummy classes appended per anchor --
            # ordinary autograd ops, so retina_cls.weight / .bias keep their reference shapes and receive their gradients
            w, b = params[-4], params[-3]
            wp = w.new_zeros((A, Kp) + tuple(w.shape[1:]))
            wp[:, :K] = w.view((A, K) + tuple(w.shape[1:]))
            bp = b.new_zeros((A, Kp))
            bp[:, :K] = b.view(A, K)
            params[-4], params[-3] = wp.view((A * Kp,) + tuple(w.shape[1:])), bp.view(A * Kp)
        fn = _ops.RetinaHeadPlanesFn if _ops.head_planes_ok(feats, params) else _ops.RetinaHeadFn
        cls, reg = fn.apply(len(feats), A, Kp, self.stacked_convs, *feats, *params)
        if Kp != K:
            cls = cls[..., :K].contiguous()
        return cls, reg

    def forward_concat(self, feats):
        return self.forward_concat_nhwc([_ops.to_nhwc(f, 'RetinaHead input') for f in feats])

    def forward_single(self, x):
        cls, reg = self.forward_concat([x])
        return cls, reg

    def forward(self, feats):
        """-> (list of per-level cls [B, HWA, K], list of per-level reg [B, HWA, 4]) -- views of the
        concatenated buffers, level order preserved."""
        cls, reg = self.forward_concat(feats)
        sizes = [int(f.shape[2]) * int(f.shape[3]) * self.num_anchors for f in feats]
        return list(cls.split(sizes, dim=1)), list(reg.split(sizes, dim=1))
