"""BiFPN neck, B200-native mirror of the reference's ``models/bifpn.py``.

``BIFPN``: five lateral 1x1 convs (+bias) then ``stack`` x ``BiFPNModule`` (reference :96-129).
``BiFPNModule``: fast-normalised weighted fusion (weights ReLU'd and normalised twice, :177-201) with
nearest x2 upsampling / 2x2 max-pooling, each node followed by a dense 3x3 conv + bias, no norm, no
activation (:151-164).  One autograd node per layer: every fusion is a single bandwidth kernel that
resamples and mixes in one pass, every conv an implicit GEMM; the reference's per-layer ``clone()`` of
all inputs (:184-186) and its ~8 element-wise launches per node disappear.
"""
import torch
import torch.nn as nn

from . import _ops
from .module import ConvModule, xavier_init


class BIFPN(nn.Module):
    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, stack=1,
                 add_extra_convs=False, extra_convs_on_inputs=True, relu_before_extra_convs=False,
                 no_norm_on_lateral=False, conv_cfg=None, norm_cfg=None, activation=None):
        super().__init__()
        assert isinstance(in_channels, list)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_ins, self.num_outs = len(in_channels), num_outs
        self.activation = activation
        self.relu_before_extra_convs = relu_before_extra_convs
        self.no_norm_on_lateral = no_norm_on_lateral
        self.stack = stack
        if end_level == -1:
            self.backbone_end_level = self.num_ins
            assert num_outs >= self.num_ins - start_level
        else:
            self.backbone_end_level = end_level
            assert end_level <= len(in_channels)
            assert num_outs == end_level - start_level
        self.start_level, self.end_level = start_level, end_level
        self.add_extra_convs, self.extra_convs_on_inputs = add_extra_convs, extra_convs_on_inputs
        if num_outs != self.backbone_end_level - start_level or add_extra_convs:
            raise NotImplementedError('extra pyramid levels (num_outs > inputs) are never built by EfficientDet '
                                      '(num_outs == num_ins == 5) and have no B200 kernel')
        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        self.stack_bifpn_convs = nn.ModuleList()
        for i in range(self.start_level, self.backbone_end_level):
            self.lateral_convs.append(ConvModule(in_channels[i], out_channels, 1, conv_cfg=conv_cfg,
                                                 norm_cfg=norm_cfg if not no_norm_on_lateral else None,
                                                 activation=self.activation, inplace=False))
        for _ in range(stack):
            self.stack_bifpn_convs.append(BiFPNModule(channels=out_channels,
                                                      levels=self.backbone_end_level - self.start_level,
                                                      conv_cfg=conv_cfg, norm_cfg=norm_cfg, activation=activation))
        self.init_weights()

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                xavier_init(m, distribution='uniform')

    def forward_nhwc(self, inputs):
        assert len(inputs) == len(self.in_channels)
        maps = [conv.forward_nhwc(inputs[i + self.start_level]) for i, conv in enumerate(self.lateral_convs)]
        for layer in self.stack_bifpn_convs:
            maps = layer.forward_nhwc(maps)
        return maps

    def forward(self, inputs):
        maps = self.forward_nhwc([_ops.to_nhwc(t, 'BIFPN input') for t in inputs])
        return tuple(_ops.to_nchw_view(t) for t in maps)


class BiFPNModule(nn.Module):
    def __init__(self, channels, levels, init=0.5, conv_cfg=None, norm_cfg=None, activation=None, eps=0.0001):
        super().__init__()
        if levels < 3:
            raise NotImplementedError('BiFPNModule needs at least 3 levels')
        self.activation, self.eps, self.levels = activation, eps, levels
        self.bifpn_convs = nn.ModuleList()
        self.w1 = nn.Parameter(torch.Tensor(2, levels).fill_(init))
        self.relu1 = nn.ReLU()
        self.w2 = nn.Parameter(torch.Tensor(3, levels - 2).fill_(init))
        self.relu2 = nn.ReLU()
        if activation is not None:
            raise NotImplementedError('BiFPN node convs carry no activation in EfficientDet')
        for _ in range(2 * (levels - 1)):
            self.bifpn_convs.append(nn.Sequential(ConvModule(channels, channels, 3, padding=1, conv_cfg=conv_cfg,
                                                             norm_cfg=norm_cfg, activation=activation,
                                                             inplace=False)))

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                xavier_init(m, distribution='uniform')

    def forward_nhwc(self, inputs):
        assert len(inputs) == self.levels
        params = []
        for seq in self.bifpn_convs:
            params += [seq[0].conv.weight, seq[0].conv.bias]
        return list(_ops.BiFPNLayerFn.apply(float(self.eps), self.levels, *inputs, self.w1, self.w2, *params))

    def forward(self, inputs):
        outs = self.forward_nhwc([_ops.to_nhwc(t, 'BiFPNModule input') for t in inputs])
        return [_ops.to_nchw_view(t) for t in outs]
