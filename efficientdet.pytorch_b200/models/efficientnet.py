"""EfficientNet backbone of the detector, B200-native.

Same public surface and state-dict keys as the reference's ``models/efficientnet.py``
(``EfficientNet.from_name/from_pretrained/get_list_features/extract_features``, ``MBConvBlock``
with ``_expand_conv/_bn0/_depthwise_conv/_bn1/_se_reduce/_se_expand/_project_conv/_bn2``), but each
MBConv block is ONE autograd node whose forward is five kernel launches
(expand GEMM+BN+swish | depthwise+BN+swish | SE mean | SE gate | project GEMM x gate+BN+skip)
instead of ~15 ATen ops (reference models/efficientnet.py:75-105).  BatchNorm layers are frozen
(eval-mode statistics, reference models/efficientdet.py:88-92) and folded into the conv epilogues.
"""
import torch
from torch import nn

from . import _ops
from .utils import (MemoryEfficientSwish, drop_connect_scale, efficientnet_params, get_model_params,
                    get_same_padding_conv2d, load_pretrained_weights, round_filters, round_repeats)


def _bn_args(bn):
    return (bn.weight, bn.bias, bn.running_mean, bn.running_var)


class MBConvBlock(nn.Module):
    """Mobile inverted bottleneck with squeeze-excite.  ``forward(inputs, drop_connect_rate)`` takes and
    returns logical NCHW tensors (channels_last memory)."""

    def __init__(self, block_args, global_params):
        super().__init__()
        self._block_args = block_args
        self._bn_mom = 1 - global_params.batch_norm_momentum
        self._bn_eps = global_params.batch_norm_epsilon
        self.has_se = (block_args.se_ratio is not None) and (0 < block_args.se_ratio <= 1)
        self.id_skip = block_args.id_skip
        if not self.has_se:
            raise NotImplementedError('the B200 MBConv kernel chain always includes squeeze-excite '
                                      '(every EfficientDet block has se_ratio=0.25)')
        Conv2d = get_same_padding_conv2d(image_size=global_params.image_size)
        inp = block_args.input_filters
        oup = inp * block_args.expand_ratio
        if block_args.expand_ratio != 1:
            self._expand_conv = Conv2d(in_channels=inp, out_channels=oup, kernel_size=1, bias=False)
            self._bn0 = nn.BatchNorm2d(num_features=oup, momentum=self._bn_mom, eps=self._bn_eps)
        k, s = block_args.kernel_size, block_args.stride
        self._depthwise_conv = Conv2d(in_channels=oup, out_channels=oup, groups=oup, kernel_size=k, stride=s,
                                      bias=False)
        self._bn1 = nn.BatchNorm2d(num_features=oup, momentum=self._bn_mom, eps=self._bn_eps)
        squeezed = max(1, int(inp * block_args.se_ratio))
        self._se_reduce = Conv2d(in_channels=oup, out_channels=squeezed, kernel_size=1)
        self._se_expand = Conv2d(in_channels=squeezed, out_channels=oup, kernel_size=1)
        self._project_conv = Conv2d(in_channels=oup, out_channels=block_args.output_filters, kernel_size=1,
                                    bias=False)
        self._bn2 = nn.BatchNorm2d(num_features=block_args.output_filters, momentum=self._bn_mom, eps=self._bn_eps)
        self._swish = MemoryEfficientSwish()

    def _has_skip(self):
        # reference models/efficientnet.py:100 -- `stride == 1` is False for the list [1] carried by the
        # first block of a stage, so only repeat blocks get the residual
        a = self._block_args
        return bool(self.id_skip and a.stride == 1 and a.input_filters == a.output_filters)

    def _kernel_cfg(self):
        a = self._block_args
        s = a.stride[0] if isinstance(a.stride, (list, tuple)) else a.stride
        pl, pr, pt, pb = self._depthwise_conv.same_pad
        return dict(k=a.kernel_size, s=s, eps=self._bn_eps, expand=a.expand_ratio != 1, skip=self._has_skip(),
                    pad_t=pt, pad_l=pl, pad_h=pt + pb, pad_w=pl + pr)

    def _params(self):
        P = []
        if self._block_args.expand_ratio != 1:
            P += [self._expand_conv.weight] + list(_bn_args(self._bn0))
        P += [self._depthwise_conv.weight] + list(_bn_args(self._bn1))
        P += [self._se_reduce.weight, self._se_reduce.bias, self._se_expand.weight, self._se_expand.bias]
        P += [self._project_conv.weight] + list(_bn_args(self._bn2))
        return P

    def forward_nhwc(self, x, drop_connect_rate=None):
        cfg = self._kernel_cfg()
        row_scale = None
        if cfg['skip'] and drop_connect_rate and self.training:
            row_scale = drop_connect_scale(x.shape[0], drop_connect_rate, x.device)
        return _ops.MBConvFn.apply(x, row_scale, cfg, *self._params())

    def forward(self, inputs, drop_connect_rate=None):
        y = self.forward_nhwc(_ops.to_nhwc(inputs, 'MBConvBlock input'), drop_connect_rate)
        return _ops.to_nchw_view(y)

    def set_swish(self, memory_efficient=True):
        """API parity; swish is always fused (and always 'memory efficient': only z is kept)."""


class EfficientNet(nn.Module):
    """Feature extractor: ``forward(images[B,3,H,W])`` -> list of the 7 stage outputs (strides 2..128)."""

    def __init__(self, blocks_args=None, global_params=None):
        super().__init__()
        assert isinstance(blocks_args, list), 'blocks_args should be a list'
        assert len(blocks_args) > 0, 'block args must be greater than 0'
        self._global_params = global_params
        self._blocks_args = blocks_args
        Conv2d = get_same_padding_conv2d(image_size=global_params.image_size)
        bn_mom = 1 - global_params.batch_norm_momentum
        bn_eps = global_params.batch_norm_epsilon
        stem = round_filters(32, global_params)
        self._conv_stem = Conv2d(3, stem, kernel_size=3, stride=2, bias=False)
        self._bn0 = nn.BatchNorm2d(num_features=stem, momentum=bn_mom, eps=bn_eps)
        self._blocks = nn.ModuleList([])
        for i in range(len(self._blocks_args)):
            a = self._blocks_args[i]._replace(
                input_filters=round_filters(self._blocks_args[i].input_filters, global_params),
                output_filters=round_filters(self._blocks_args[i].output_filters, global_params),
                num_repeat=round_repeats(self._blocks_args[i].num_repeat, global_params))
            self._blocks_args[i] = a
            self._blocks.append(MBConvBlock(a, global_params))
            if a.num_repeat > 1:
                a = a._replace(input_filters=a.output_filters, stride=1)
                self._blocks_args[i] = a
            for _ in range(a.num_repeat - 1):
                self._blocks.append(MBConvBlock(a, global_params))
        # classifier tail: never used by the detector but part of the checkpoint schema
        last = self._blocks_args[-1].output_filters
        head = round_filters(1280, global_params)
        self._conv_head = Conv2d(last, head, kernel_size=1, bias=False)
        self._bn1 = nn.BatchNorm2d(num_features=head, momentum=bn_mom, eps=bn_eps)
        self._avg_pooling = nn.AdaptiveAvgPool2d(1)
        self._dropout = nn.Dropout(global_params.dropout_rate)
        self._fc = nn.Linear(head, global_params.num_classes)
        self._swish = MemoryEfficientSwish()

    def set_swish(self, memory_efficient=True):
        for b in self._blocks:
            b.set_swish(memory_efficient)

    def extract_features_nhwc(self, inputs):
        _ops.check_cuda_f32(inputs, 'EfficientNet input')
        bn = self._bn0
        x = _ops.StemFn.apply(inputs, self._conv_stem.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                              bn.eps)
        feats = []
        stage, seen = 0, 0
        nblocks = len(self._blocks)
        for idx, block in enumerate(self._blocks):
            rate = self._global_params.drop_connect_rate
            if rate:
                rate *= float(idx) / nblocks
            x = block.forward_nhwc(x, drop_connect_rate=rate)
            seen += 1
            if seen == self._blocks_args[stage].num_repeat:
                seen = 0
                stage += 1
                feats.append(x)
        return feats

    def extract_features(self, inputs):
        return [_ops.to_nchw_view(f) for f in self.extract_features_nhwc(inputs)]

    def forward(self, inputs):
        return self.extract_features(inputs)

    @classmethod
    def from_name(cls, model_name, override_params=None):
        cls._check_model_name_is_valid(model_name)
        blocks_args, global_params = get_model_params(model_name, override_params)
        return cls(blocks_args, global_params)

    @classmethod
    def from_pretrained(cls, model_name, num_classes=1000):
        model = cls.from_name(model_name, override_params={'num_classes': num_classes})
        load_pretrained_weights(model, model_name, load_fc=(num_classes == 1000))
        return model

    @classmethod
    def get_image_size(cls, model_name):
        cls._check_model_name_is_valid(model_name)
        return efficientnet_params(model_name)[2]

    @classmethod
    def _check_model_name_is_valid(cls, model_name, also_need_pretrained_weights=False):
        n = 4 if also_need_pretrained_weights else 8
        valid = ['efficientnet-b' + str(i) for i in range(n)]
        if model_name not in valid:
            raise ValueError('model_name should be one of: ' + ', '.join(valid))

    def get_list_features(self):
        return [a.output_filters for a in self._blocks_args]
